// valid_main.cpp -- drop-in `valid` command line (the reference's external-validation tool, scr/main_valid.cpp +
// scr/validate.cpp): per LD block, nume = z1'z2 and deno = z1' (X'X/n) z1 over the SNPs shared by a
// DBSLMM result file, an external summary file and the reference panel, written to <r2>.txt as "nume deno" lines.
// Same options, file formats, matching rules and quirks as the reference; the genotype work (readSNPIm + nomalizeVec
// per SNP, the n x m product, validate.cpp:243-256) runs on the GPU through the C ABI: the decoder and the exact
// integer Gram of the fit with tau = 1, followed by a quadratic-form kernel (fit_args.quadform_out).  No CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "dbslmm_b200.h"
#include "ingest.hpp"

using namespace std;
using namespace dbslmm_host;

namespace {

struct Param { string d, s, r, b, r2; double mafMax = 0.0; };        // PARAM, validate.hpp:28-37 (mafMax is uninitialised there)

void print_header() {                                                 // VALID::printHeader, validate.cpp:38-46
    cout << endl;
    cout << "*************************************************************" << endl;
    cout << "  External validation of DBSLMM (valid), B200 build" << endl;
    cout << "  Type ./valid -h for detailed help" << endl;
    cout << "*************************************************************" << endl;
    cout << endl;
}
void print_help() {                                                   // VALID::printHelp, validate.cpp:48-60
    cout << " FILE I/O RELATED OPTIONS" << endl;
    cout << " -d        [filename]  " << " specify input the DBSLMM result file." << endl;
    cout << " -s        [filename]  " << " specify input the external summary statistics file (snp a1 maf z)." << endl;
    cout << " -r        [filename]  " << " specify input the bfile of reference data." << endl;
    cout << " -mafMax   [num]       " << " specify input the maximium of the difference between reference panel and summary data." << endl;
    cout << " -b        [filename]  " << " specify input the block information." << endl;
    cout << " -r2       [filename]  " << " specify output the prefix of the numerator / denominator file." << endl;
}

// VALID::Assign, validate.cpp:62-118: a value that is missing or begins with '-' is ignored
void assign(int argc, char** argv, Param& p) {
    auto is = [&](int i, const char* a, const char* b) { return strcmp(argv[i], a) == 0 || strcmp(argv[i], b) == 0; };
    for (int i = 0; i < argc; i++) {
        string* dst = nullptr;
        bool num = false;
        if (is(i, "--dbslmm", "-d")) dst = &p.d;
        else if (is(i, "--summ", "-s")) dst = &p.s;
        else if (is(i, "--reference", "-r")) dst = &p.r;
        else if (is(i, "--mafMax", "-mafMax")) num = true;
        else if (is(i, "--block", "-b")) dst = &p.b;
        else if (is(i, "--R2", "-r2")) dst = &p.r2;
        else continue;
        if (i + 1 >= argc || argv[i + 1] == nullptr || argv[i + 1][0] == '-') continue;
        ++i;
        if (num) p.mafMax = atof(argv[i]); else *dst = argv[i];
    }
}

struct Summs { string snp, a1; double maf = 0.0, z = 0.0; };          // SUMMS, dtpr.hpp:74-80
struct Summc { string snp, a1; double maf, z1, z2; };                 // SUMMC, dtpr.hpp:83-90
struct Summp { string snp; double z1, z2; long pos, ps; };            // SUMMP, dtpr.hpp:93-100
struct Alleleb { long pos, ps; string a1, a2; double maf; };          // ALLELEB, dtpr.hpp:43-50

void split(const string& line, char sep, vector<string>& out) {
    out.clear();
    stringstream ss(line);
    string el;
    while (getline(ss, el, sep)) out.push_back(el);
}

}  // namespace

int main(int argc, char* argv[]) {
    Param cPar;
    if (argc <= 1) { print_header(); return EXIT_SUCCESS; }
    if (argc == 2 && argv[1][0] == '-' && argv[1][1] == 'h') { print_help(); return EXIT_SUCCESS; }
    assign(argc, argv, cPar);

    // ---- VALID::BatchRun, validate.cpp:120-265
    cout << "Options: " << endl;
    cout << "-d:      " << cPar.d << endl;
    cout << "-s:      " << cPar.s << endl;
    cout << "-r:      " << cPar.r << endl;
    cout << "-mafMax: " << cPar.mafMax << endl;
    cout << "-b:      " << cPar.b << endl;
    cout << "-r2:      " << cPar.r2 << endl;
    const string ref_fam_str = cPar.r + ".fam";
    if (cPar.d.empty()) { cerr << "ERROR: -d is no parameter!" << endl; exit(1); }
    if (cPar.s.empty()) { cerr << "ERROR: -s is no parameter!" << endl; exit(1); }
    if (cPar.r.empty()) { cerr << "ERROR: -r is no parameter!" << endl; exit(1); }
    {
        ifstream dS(cPar.d.c_str()), sS(cPar.s.c_str()), rS(ref_fam_str.c_str()), bS(cPar.b.c_str());
        if (!dS) { cerr << "ERROR: " << cPar.d << " dose not exist!" << endl; exit(1); }
        if (!sS) { cerr << "ERROR: " << cPar.s << " dose not exist!" << endl; exit(1); }
        if (!rS) { cerr << "ERROR: " << cPar.r << " dose not exist!" << endl; exit(1); }
        if (!bS) { cerr << "ERROR: " << cPar.b << " dose not exist!" << endl; exit(1); }
    }
    cout << "Reading PLINK FAM file from [" << cPar.r << ".fam]" << endl;
    const int n_ref = get_row(ref_fam_str);
    cout << n_ref << " individuals to be included from reference FAM file." << endl;

    vector<string> t;
    string line;
    // readDBSLMM, dtpr.cpp:223-245: space-separated, columns snp a1 z(=column 3)
    cout << "Reading DBSLMM result file from [" << cPar.d << "]" << endl;
    vector<Summs> dbslmm;
    {
        ifstream f(cPar.d.c_str());
        while (getline(f, line)) {
            split(line, ' ', t);
            if (t.size() < 3) continue;
            Summs s; s.snp = t[0]; s.a1 = t[1]; s.z = atof(t[2].c_str());
            dbslmm.push_back(s);
        }
    }
    cout << dbslmm.size() << " SNPs in DBSLMM result. " << endl;
    // readExt, dtpr.cpp:248-270: space-separated, columns snp a1 maf z; first occurrence of an rs id wins (map::insert)
    cout << "Reading summary statistics file from [" << cPar.s << "]" << endl;
    map<string, Summs> summ_ext;
    {
        ifstream f(cPar.s.c_str());
        while (getline(f, line)) {
            split(line, ' ', t);
            if (t.size() < 4) continue;
            Summs s; s.snp = t[0]; s.a1 = t[1]; s.maf = atof(t[2].c_str()); s.z = atof(t[3].c_str());
            summ_ext.insert(make_pair(t[0], s));
        }
    }
    cout << summ_ext.size() << " SNPs in external result. " << endl;
    // matchSumm, dtpr.cpp:411-434: z2 changes sign when the alleles differ
    vector<Summc> summ_comb;
    {
        int dis_count = 0;
        for (const Summs& d : dbslmm) {
            auto it = summ_ext.find(d.snp);
            if (it == summ_ext.end()) continue;
            const bool same = it->second.a1 == d.a1;
            if (!same) dis_count++;
            summ_comb.push_back(Summc{it->second.snp, d.a1, it->second.maf, d.z, same ? it->second.z : -it->second.z});
        }
        cout << "Number of allele discrepency: " << dis_count << endl;
    }
    cout << summ_comb.size() << " SNPs are intersection of DBSLMM and external summary statistics. " << endl;

    // ---- reference panel on the GPU; its statistics kernel is the MAF pre-pass of readBim (dtpr.cpp:136-146)
    cout << "Reading PLINK BIM file from [" << cPar.r << ".bim]" << endl;
    const bool constr = !(fabs(cPar.mafMax - 1.0) < 1e-10);
    const int64_t n_snp = get_row(cPar.r + ".bim");
    vector<uint8_t> bed;
    if (!read_bed(cPar.r + ".bed", n_snp, n_ref, bed)) { cerr << "ERROR: cannot read SNP-major " << cPar.r << ".bed" << endl; exit(1); }
    dbslmm_b200_handle* h = nullptr;
    if (dbslmm_b200_create(0, &h) != DBSLMM_B200_OK) { cerr << "ERROR: cannot initialise the GPU (there is no CPU fallback)" << endl; exit(2); }
    if (dbslmm_b200_load_bed(h, bed.data(), n_snp, n_ref) != DBSLMM_B200_OK) { cerr << "ERROR: load_bed: " << dbslmm_b200_last_error(h) << endl; exit(2); }
    vector<double> maf((size_t)n_snp, 0.0);
    if (constr) {
        cout << "Calculating MAF of reference panel ..." << endl;
        dbslmm_b200_snp_stats(h, maf.data(), nullptr);
    } else {
        cout << "[WARNING] Do not consider the difference between reference panel and external summary data ..." << endl;
    }
    map<string, Alleleb> ref_bim;
    {
        ifstream f((cPar.r + ".bim").c_str());
        long count = 0;
        while (getline(f, line)) {
            split(line, '\t', t);
            if (t.size() >= 6) ref_bim.insert(make_pair(t[1], Alleleb{count, (long)atoi(t[3].c_str()), t[4], t[5], maf[(size_t)count]}));
            count++;
        }
    }
    cout << ref_bim.size() << " SNPs to be included from reference BIM file." << endl;
    // matchAll, dtpr.cpp:436-453 (an rs id missing from the .bim has empty alleles there, so it never matches)
    vector<Summp> summ_comb_r;
    for (const Summc& c : summ_comb) {
        auto it = ref_bim.find(c.snp);
        if (it == ref_bim.end()) continue;
        if (it->second.a1 == c.a1 && fabs(it->second.maf - c.maf) < cPar.mafMax)
            summ_comb_r.push_back(Summp{c.snp, c.z1, c.z2, it->second.pos, it->second.ps});
    }
    stable_sort(summ_comb_r.begin(), summ_comb_r.end(), [](const Summp& a, const Summp& b) { return a.ps < b.ps; });
    cout << summ_comb_r.size() << " SNPs intersect." << endl;

    vector<Block> block_dat;
    read_block(cPar.b, block_dat);
    const int num_block = (int)block_dat.size();
    cout << num_block << " blocks for the chromesome." << endl;

    // ---- block loop, validate.cpp:225-259: a sequential scan that stops at the first SNP outside [start, end) -- SNPs
    // in a gap before a block stall it for good, as in the reference
    vector<int32_t> off(num_block + 1, 0), pos;
    vector<double> z1, nume(num_block, 0.0), deno(num_block, 0.0);
    size_t cur = 0;
    for (int b = 0; b < num_block; ++b) {
        for (; cur < summ_comb_r.size(); ++cur) {
            const Summp& s = summ_comb_r[cur];
            if (!(s.ps >= block_dat[b].start && s.ps < block_dat[b].end)) break;
            pos.push_back((int32_t)s.pos);
            z1.push_back(s.z1);
            nume[b] += s.z1 * s.z2;                                            // :257
        }
        off[b + 1] = (int32_t)pos.size();
    }
    if (num_block > 0) {
        dbslmm_b200_fit_args a{};
        a.n_blocks = num_block;
        a.s_off = off.data(); a.s_pos = pos.data(); a.s_z = z1.data();
        a.tau = 1.0;                                                           // Sigma = X'X / n, no shrinkage (:255-256)
        a.solver = DBSLMM_B200_SOLVER_CHOLESKY;
        a.quadform_out = deno.data();                                          // :258
        const int rc = dbslmm_b200_fit(h, &a);
        if (rc < 0) { cerr << "ERROR: " << dbslmm_b200_last_error(h) << endl; exit(2); }
    }
    dbslmm_b200_destroy(h);
    const string r2_str = cPar.r2 + ".txt";
    ofstream out(r2_str.c_str());
    for (int i = 0; i < num_block; ++i) out << nume[i] << " " << deno[i] << endl;
    return EXIT_SUCCESS;
}
