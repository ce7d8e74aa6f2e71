// dbslmm_main.cpp -- the `dbslmm` command line on top of the B200 C ABI.
//
// Drop-in for the reference's CLI (scr/main_dbslmm.cpp:31-48, DBSLMM::Assign scr/dbslmm.cpp:67-172,
// DBSLMM::BatchRun :174-398): same options and aliases, same input formats, same `<eff>.txt`
// and `<eff>.badsnps` outputs.  The block fit itself (DBSLMMFIT::est) runs on the GPU(s)
// through include/dbslmm_b200.h; there is no CPU path.
//
// Differences on purpose:
//   * -test_indicator_file / -dat_str are OPTIONAL (the fork crashes without them): when both are given the
//     asymptotic-variance side channel runs on the GPU and `variance.txt` (n_test x num_block, Armadillo
//     ASCII layout) is written to the working directory like scr/dbslmmfit.cpp:242; a SNP that is absent
//     from the test .bim is an error (the reference silently misaligns its vectors in that case).
//   * hidden additions: --solver chol|pcg, --gpus N, --tau T, --h2-folds a,b,c (one Gram, several
//     ridge folds, outputs <eff>_f<k>.txt), --dump-beta-bin FILE (FP64 betas; the text output
//     only has 6 significant digits), --verbose.
//   * --manifest FILE: genome-wide in one process.  The reference is run once per chromosome
//     by a bash loop (DBSLMM_script.sh:69,118); here each manifest line `s<TAB>l<TAB>r<TAB>b<TAB>eff` (l may be
//     `-`) is one chromosome, all panels are concatenated into one GPU plan and every chromosome gets its own
//     `<eff>.txt` / `<eff>.badsnps` with the same rows as a separate run (values agree to FP64 rounding: the
//     split-K slicing of the Cholesky depends on how many blocks share a panel step).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <sys/time.h>
#include <thread>
#include <unordered_map>
#include <vector>

#include "dbslmm_b200.h"
#include "ingest.hpp"

using namespace std;
using namespace dbslmm_host;

struct Param {                                   // PARAM, scr/dbslmm.hpp:29-45 (+ explicit defaults)
    string s, l, r, b, eff, test_indicator_file, dat_str;
    int n = 0, nsnp = 0, t = 1;
    double mafMax = 1.0, h = 0.0;
    // additions
    string solver = "chol", dump_bin;
    int gpus = 1;
    double tau = 0.8;
    vector<double> folds;
    bool verbose = false;
    string manifest;                             // --manifest: several chromosomes (s l r b eff per line) in ONE run / GPU plan
};

static void print_header() {
    cout << endl;
    cout << "*************************************************************" << endl;
    cout << "  Deterministic Bayesian Sparse Linear Mixed Model (DBSLMM)  " << endl;
    cout << "  B200-native block fit (dbslmm_b200), CLI-compatible with   " << endl;
    cout << "  DBSLMM 0.3                                                 " << endl;
    cout << "  GNU General Public License                                 " << endl;
    cout << "  For Help, Type ./dbslmm -h                                 " << endl;
    cout << "*************************************************************" << endl;
    cout << endl;
}

static void print_help() {
    cout << " FILE I/O RELATED OPTIONS" << endl;
    cout << " -s        [filename]  " << " specify input the summary data for the small effect SNPs." << endl;
    cout << " -l        [filename]  " << " specify input the summary data for the large effect SNPs." << endl;
    cout << " -r        [filename]  " << " specify input the bfile of reference data." << endl;
    cout << " -n        [num]       " << " specify input the sample size of the summary data." << endl;
    cout << " -mafMax   [num]       " << " specify input the maximium of the difference between reference panel and summary data." << endl;
    cout << " -nsnp     [num]  " << " specify input the number of snp." << endl;
    cout << " -b        [num]       " << " specify input the block information." << endl;
    cout << " -h        [num]       " << " specify input the heritability." << endl;
    cout << " -t        [filename]  " << " specify input thread." << endl;
    cout << " -eff      [filename]  " << " specify output the estimate effect SNPs." << endl;
}

static bool opt(const char* a, const char* lng, const char* sht) { return strcmp(a, lng) == 0 || strcmp(a, sht) == 0; }

static void assign(int argc, char** argv, Param& p) {       // scr/dbslmm.cpp:67-172
    for (int i = 0; i < argc; i++) {
        auto val = [&](string& dst) -> bool {
            if (argv[i + 1] == NULL || argv[i + 1][0] == '-') return false;   // :74 values starting with '-' are skipped
            ++i;
            dst.assign(argv[i]);
            return true;
        };
        string v;
        if (opt(argv[i], "--smallEff", "-s")) { if (val(v)) p.s = v; }
        else if (opt(argv[i], "--largeEff", "-l")) { if (val(v)) p.l = v; }
        else if (opt(argv[i], "--reference", "-r")) { if (val(v)) p.r = v; }
        else if (opt(argv[i], "--N", "-n")) { if (val(v)) p.n = atoi(v.c_str()); }
        else if (opt(argv[i], "--mafMax", "-mafMax")) { if (val(v)) p.mafMax = atof(v.c_str()); }
        else if (opt(argv[i], "--numSNP", "-nsnp")) { if (val(v)) p.nsnp = atoi(v.c_str()); }
        else if (opt(argv[i], "--block", "-b")) { if (val(v)) p.b = v; }
        else if (opt(argv[i], "--Heritability", "-h")) { if (val(v)) p.h = atof(v.c_str()); }
        else if (opt(argv[i], "--Thread", "-t")) { if (val(v)) p.t = atoi(v.c_str()); }
        else if (opt(argv[i], "--EFF", "-eff")) { if (val(v)) p.eff = v; }
        else if (opt(argv[i], "--test_indicator_file", "-test_indicator_file")) { if (val(v)) p.test_indicator_file = v; }
        else if (opt(argv[i], "--dat_str", "-dat_str")) { if (val(v)) p.dat_str = v; }
        else if (opt(argv[i], "--solver", "-solver")) { if (val(v)) p.solver = v; }
        else if (opt(argv[i], "--gpus", "-gpus")) { if (val(v)) p.gpus = atoi(v.c_str()); }
        else if (opt(argv[i], "--tau", "-tau")) { if (val(v)) p.tau = atof(v.c_str()); }
        else if (opt(argv[i], "--dump-beta-bin", "-dump-beta-bin")) { if (val(v)) p.dump_bin = v; }
        else if (opt(argv[i], "--h2-folds", "-h2-folds")) {
            if (val(v)) { stringstream ss(v); string e; while (getline(ss, e, ',')) p.folds.push_back(atof(e.c_str())); }
        }
        else if (opt(argv[i], "--manifest", "-manifest")) { if (val(v)) p.manifest = v; }
        else if (opt(argv[i], "--verbose", "-verbose")) p.verbose = true;
    }
}

static double walltime() { struct timeval t; gettimeofday(&t, NULL); return (double)t.tv_sec + (double)t.tv_usec * 1e-6; }

// fn(i) for i in [0, n) on up to `width` host threads (chromosomes are independent until the block plan is merged)
template <class F>
static void parallel_for(int n, int width, F&& fn) {
    width = std::max(1, std::min(width, n));
    if (width == 1) { for (int i = 0; i < n; ++i) fn(i); return; }
    std::atomic<int> next{0};
    vector<thread> th;
    for (int t = 0; t < width; ++t)
        th.emplace_back([&]() { for (int i = next.fetch_add(1); i < n; i = next.fetch_add(1)) fn(i); });
    for (auto& x : th) x.join();
}

struct Shard {                                   // what one GPU fits
    vector<int> blocks;                          // global block ids, ascending
    vector<int32_t> s_off, s_pos, l_off, l_pos;
    vector<double> s_z, l_z, beta_s, beta_l;
    vector<int32_t> status, s_tpos, l_tpos;
    vector<double> variance;                     // [n_folds][blocks][n_test]
    vector<uint8_t> bed;                         // compact .bed (only when gpus > 1)
    int64_t n_rows = 0;
    int rc = 0;
    string err;
    dbslmm_b200_timing timing{};
};

// one chromosome's worth of inputs (the reference processes exactly one per process)
struct Job {
    string s, l, r, b, eff;
    int64_t n_snp_ref = 0, row0 = 0;             // rows of this panel inside the concatenated .bed
    int num_block = 0, block0 = 0;
    BimMap bim;
    Info info_s, info_l;
    bool with_large = false;
    int n_fam = 0;
    string log;                                  // this chromosome's share of the stdout log (printed in manifest order)
    string err;                                  // first fatal problem met by a worker thread
};

int main(int argc, char* argv[]) {
    if (argc <= 1) { print_header(); return EXIT_SUCCESS; }                                   // main_dbslmm.cpp:37-40
    if (argc == 2 && argv[1][0] == '-' && argv[1][1] == 'h') { print_help(); return EXIT_SUCCESS; }   // :41-44
    Param cPar;
    assign(argc, argv, cPar);

    vector<Job> jobs;
    if (cPar.manifest.empty()) {
        Job j; j.s = cPar.s; j.l = cPar.l; j.r = cPar.r; j.b = cPar.b; j.eff = cPar.eff;
        jobs.push_back(j);
    } else {
        ifstream mf(cPar.manifest.c_str());
        if (!mf) { cerr << "ERROR: " << cPar.manifest << " dose not exist!" << endl; exit(1); }
        string line;
        while (getline(mf, line)) {
            if (line.empty() || line[0] == '#') continue;
            vector<string> t; stringstream ss(line); string e;
            while (getline(ss, e, '\t')) t.push_back(e);
            if (t.size() < 5) { cerr << "ERROR: manifest line needs s, l, r, b, eff (tab separated)" << endl; exit(1); }
            Job j; j.s = t[0]; j.l = (t[1] == "-" ? string() : t[1]); j.r = t[2]; j.b = t[3]; j.eff = t[4];
            jobs.push_back(j);
        }
        if (jobs.empty()) { cerr << "ERROR: empty manifest" << endl; exit(1); }
    }

    cout << "Options: " << endl;                                                              // dbslmm.cpp:181-192
    cout << "-s:      " << jobs[0].s << endl;
    cout << "-l:      " << jobs[0].l << endl;
    cout << "-r:      " << jobs[0].r << endl;
    cout << "-nsnp:   " << cPar.nsnp << endl;
    cout << "-n:      " << cPar.n << endl;
    cout << "-mafMax: " << cPar.mafMax << endl;
    cout << "-b:      " << jobs[0].b << endl;
    cout << "-h:      " << cPar.h << endl;
    cout << "-t:      " << cPar.t << endl;
    cout << "-eff:    " << jobs[0].eff << endl;
    cout << "-test_indicator_file:  " << cPar.test_indicator_file << endl;
    if (jobs.size() > 1) cout << "--manifest: " << jobs.size() << " chromosomes in one run" << endl;

    for (const Job& j : jobs) {                                                               // :194-219
        const string ref_fam = j.r + ".fam";
        ifstream seff(j.s.c_str()), reff(ref_fam.c_str()), beff(j.b.c_str());
        if (j.s.size() == 0) { cerr << "ERROR: -s is no parameter!" << endl; exit(1); }
        if (!beff) { cerr << "ERROR: " << j.b << " dose not exist!" << endl; exit(1); }
        if (!seff) { cerr << "ERROR: " << j.s << " dose not exist!" << endl; exit(1); }
        if (!reff) { cerr << "ERROR: " << j.r << " dose not exist!" << endl; exit(1); }
        if (j.b.size() == 0) { cerr << "ERROR: -b is no parameter!" << endl; exit(1); }
        if (j.r.size() == 0) { cerr << "ERROR: " << j.r << " dose not exist!" << endl; exit(1); }
    }
    if (cPar.h > 1 || cPar.h < 0) { cerr << "ERROR: -h is not correct (0, 1)!" << endl; exit(1); }   // :220-227
    // the reference accepts -h 0 and then prints NaN effects (the ridge 1 / (h / nsnp * n) is infinite); say so instead
    if (cPar.h == 0) { cerr << "ERROR: -h 0 leaves the ridge 1/(h/nsnp*n) undefined (the reference prints NaN effects): give -h > 0" << endl; exit(1); }
    if (cPar.t > 100 || cPar.t < 1) { cerr << "ERROR: -t is not correct (1, 100)!" << endl; exit(1); }
    if (cPar.n <= 0 || cPar.nsnp <= 0) { cerr << "ERROR: -n and -nsnp must be positive!" << endl; exit(1); }
    const int solver = (cPar.solver == "pcg") ? DBSLMM_B200_SOLVER_PCG : DBSLMM_B200_SOLVER_CHOLESKY;
    const bool want_var = !cPar.test_indicator_file.empty() && !cPar.dat_str.empty();
    if (want_var && jobs.size() > 1) { cerr << "ERROR: the variance side channel (-dat_str) is per chromosome: not available with --manifest" << endl; exit(1); }
    if (want_var && solver != DBSLMM_B200_SOLVER_CHOLESKY) { cerr << "ERROR: the variance side channel needs --solver chol" << endl; exit(1); }

    const int n_dev = dbslmm_b200_device_count();
    if (n_dev <= 0) { cerr << "ERROR: no CUDA device: dbslmm_b200 has no CPU fallback." << endl; exit(2); }
    const int n_gpus = std::max(1, std::min(cPar.gpus, n_dev));

    // host threads for the text / file phases: -t as given for one chromosome; a manifest run takes the machine
    const int host_threads = jobs.size() > 1 ? std::max(cPar.t, (int)std::thread::hardware_concurrency()) : cPar.t;
    const double t_start = walltime();
    // GPU contexts come up in the background while the text files are read (~0.3 s each on a cold process)
    const bool constr = !(fabs(cPar.mafMax - 1.0) < 1e-10);                                   // :238-241
    vector<dbslmm_b200_handle*> hs(n_gpus, nullptr);
    vector<int> hs_rc(n_gpus, 0);
    vector<thread> gpu_init;
    for (int g = 0; g < n_gpus; ++g) gpu_init.emplace_back([&, g]() { hs_rc[g] = dbslmm_b200_create(g, &hs[g]); });

    // ---- reference panels: .fam / .bim of every chromosome (in parallel), then the .bed files straight into ONE
    // page-locked buffer, concatenated row-wise
    parallel_for((int)jobs.size(), host_threads, [&](int i) {
        Job& j = jobs[i];
        ostringstream out;
        out << "Reading reference PLINK FAM file from [" << j.r << ".fam]" << endl;
        j.n_fam = get_row(j.r + ".fam");                                                      // :232
        out << j.n_fam << " individuals to be included from reference FAM file." << endl;
        out << "Reading reference PLINK BIM file from [" << j.r << ".bim]" << endl;
        j.n_snp_ref = read_bim(j.r + ".bim", j.bim);
        out << j.bim.size() << " SNPs to be included from reference BIM file." << endl;
        j.log = out.str();
    });
    int n_ref = -1;
    int64_t n_snp_all = 0;
    for (Job& j : jobs) {
        if (n_ref >= 0 && j.n_fam != n_ref) { cerr << "ERROR: all panels of a manifest must hold the same individuals" << endl; exit(1); }
        n_ref = j.n_fam;
        j.row0 = n_snp_all;
        n_snp_all += j.n_snp_ref;
    }
    for (auto& t : gpu_init) t.join();
    for (int g = 0; g < n_gpus; ++g)
        if (hs_rc[g] != DBSLMM_B200_OK) { cerr << "ERROR: cannot initialise GPU " << g << endl; exit(2); }
    const size_t pitch_all = (size_t)((n_ref + 3) / 4);
    uint8_t* bed = nullptr;
    {
        void* p = nullptr;
        if (dbslmm_b200_host_alloc(hs[0], (uint64_t)n_snp_all * pitch_all + 64, &p) != DBSLMM_B200_OK) {
            cerr << "ERROR: cannot allocate " << n_snp_all * pitch_all << " bytes of page-locked host memory: " << dbslmm_b200_last_error(hs[0]) << endl; exit(2);
        }
        bed = (uint8_t*)p;
    }
    parallel_for((int)jobs.size(), host_threads, [&](int i) {
        Job& j = jobs[i];
        if (!read_bed_into(j.r + ".bed", j.n_snp_ref, n_ref, bed + (size_t)j.row0 * pitch_all)) j.err = "ERROR: cannot read SNP-major " + j.r + ".bed";
    });
    for (Job& j : jobs) if (!j.err.empty()) { cerr << j.err << endl; exit(1); }
    const double t_panel = walltime();
    // With a MAF constraint GPU 0 takes the whole panel now: its statistics kernel IS the MAF pre-pass
    // (dtpr.cpp:93-102).  Without one (mafMax == 1, dbslmm.cpp:238-241) nothing needs the panel before the fit, and
    // it travels with dbslmm_b200_fit (fit_args.bed): the upload then overlaps the fit.
    const bool early_load = constr;
    if (early_load && dbslmm_b200_load_bed(hs[0], bed, n_snp_all, n_ref) != DBSLMM_B200_OK) {
        cerr << "ERROR: load_bed: " << dbslmm_b200_last_error(hs[0]) << endl; exit(2);
    }
    for (Job& j : jobs) { cout << j.log; j.log.clear(); }
    vector<double> ref_maf;
    if (constr) {
        cout << "Calculating MAF of reference panel ..." << endl;
        ref_maf.resize((size_t)n_snp_all);
        if (dbslmm_b200_snp_stats(hs[0], ref_maf.data(), nullptr) != DBSLMM_B200_OK) {
            cerr << "ERROR: snp_stats: " << dbslmm_b200_last_error(hs[0]) << endl; exit(2);
        }
    } else {
        cout << "[WARNING] Do not consider the difference between reference panel and summary data ..." << endl;
    }

    // ---- per chromosome (in parallel): blocks, summary statistics, matching, block assignment, badsnps
    parallel_for((int)jobs.size(), host_threads, [&](int ji) {
        Job& j = jobs[ji];
        ostringstream out;
        vector<Block> block_dat;
        read_block(j.b, block_dat);                                                           // :248
        j.num_block = (int)block_dat.size();
        const double* maf = constr ? ref_maf.data() + j.row0 : nullptr;
        out << "Reading summary data of small effect SNPs from [" << j.s << "]" << endl;
        Summ summ_s;
        read_summ(j.s, summ_s);
        Info inter_s;
        vector<char> matched_s;
        int dis = 0, mafc = 0;
        match_ref(summ_s, j.bim, maf, cPar.mafMax, inter_s, matched_s, dis, mafc);
        out << "Number of allele discrepency: " << dis << endl;
        out << "Number of maf discrepency:    " << mafc << endl;
        out << "After filtering, " << inter_s.size() << " small effect SNPs are selected." << endl;
        add_block(inter_s, block_dat, j.info_s);
        const string badsnps_str = j.eff + ".badsnps";
        string bad;
        for (size_t i = 0; i < summ_s.size(); ++i)
            if (!matched_s[i]) { bad += summ_s.snp[i]; bad += " 0\n"; }                      // :282-285
        ifstream leff(j.l.c_str());
        Info inter_l;
        if (!j.l.empty() && leff) {                                                           // :292-317
            out << "Reading summary data of large effect SNPs from [" << j.l << "]" << endl;
            Summ summ_l;
            read_summ(j.l, summ_l);
            vector<char> matched_l;
            match_ref(summ_l, j.bim, maf, cPar.mafMax, inter_l, matched_l, dis, mafc);
            out << "Number of allele discrepency: " << dis << endl;
            out << "Number of maf discrepency:    " << mafc << endl;
            if (inter_l.size() != 0) {
                add_block(inter_l, block_dat, j.info_l);
                out << "After filtering, " << inter_l.size() << " large effect SNPs are selected." << endl;
            } else {
                out << "After filtering, no large effect SNP is selected." << endl;
            }
            for (size_t i = 0; i < summ_l.size(); ++i)
                if (!matched_l[i]) { bad += summ_l.snp[i]; bad += " 1\n"; }
        }
        ofstream badsnpsFout(badsnps_str.c_str(), ios::binary);
        badsnpsFout.write(bad.data(), (streamsize)bad.size());
        badsnpsFout.close();
        j.with_large = inter_l.size() != 0;                                                   // :325 / :366
        BimMap().swap(j.bim);                                                                 // not needed any more
        j.log = out.str();
    });
    int num_block = 0;
    for (Job& j : jobs) {
        cout << j.log; j.log.clear();
        j.block0 = num_block;
        num_block += j.num_block;
    }
    const double t_ingest = walltime();
    bool with_large = false;
    for (const Job& j : jobs) with_large = with_large || j.with_large;

    // ---- ONE block plan over all chromosomes: CSR offsets (global block ids), shard over GPUs
    vector<int32_t> s_off(1, 0), l_off(1, 0), s_pos_all, l_pos_all;
    vector<double> s_z_all, l_z_all;
    for (const Job& j : jobs) {
        const vector<int32_t> so = block_offsets(j.info_s, j.num_block);
        const vector<int32_t> lo = j.with_large ? block_offsets(j.info_l, j.num_block) : vector<int32_t>((size_t)j.num_block + 1, 0);
        const int32_t sb = s_off.back(), lb = l_off.back();
        for (int b = 1; b <= j.num_block; ++b) { s_off.push_back(sb + so[b]); l_off.push_back(lb + lo[b]); }
        for (size_t i = 0; i < j.info_s.size(); ++i) { s_pos_all.push_back((int32_t)(j.info_s.pos[i] + j.row0)); s_z_all.push_back(j.info_s.z[i]); }
        if (j.with_large)
            for (size_t i = 0; i < j.info_l.size(); ++i) { l_pos_all.push_back((int32_t)(j.info_l.pos[i] + j.row0)); l_z_all.push_back(j.info_l.z[i]); }
    }
    vector<int32_t> m_s(num_block), m_l(num_block, 0), owner(num_block, 0);
    for (int b = 0; b < num_block; ++b) {
        m_s[b] = s_off[b + 1] - s_off[b];
        m_l[b] = l_off[b + 1] - l_off[b];
    }
    if (n_gpus > 1) dbslmm_b200_plan_shards(num_block, m_s.data(), m_l.data(), n_ref, n_gpus, owner.data(), nullptr);

    vector<double> folds = cPar.folds;
    if (folds.empty()) folds.push_back(1.0);
    vector<double> sigma_s(folds.size());
    for (size_t f = 0; f < folds.size(); ++f) sigma_s[f] = folds[f] * cPar.h / (double)cPar.nsnp;   // :332 (x fold)
    const int n_folds = (int)folds.size();

    vector<Shard> shards(n_gpus);
    for (int g = 0; g < n_gpus; ++g) {
        Shard& sh = shards[g];
        sh.s_off.push_back(0);
        if (with_large) sh.l_off.push_back(0);
        // several GPUs: every GPU is handed the whole host panel and uploads only the rows of ITS blocks
        // (DBSLMM_B200_FLAG_PANEL_SUBSET): no per-GPU compaction on the host
        auto map_row = [&](int32_t p) -> int32_t { return p; };
        for (int b = 0; b < num_block; ++b) {
            if (owner[b] != g) continue;
            sh.blocks.push_back(b);
            for (int j = s_off[b]; j < s_off[b + 1]; ++j) { sh.s_pos.push_back(map_row(s_pos_all[j])); sh.s_z.push_back(s_z_all[j]); }
            sh.s_off.push_back((int32_t)sh.s_pos.size());
            if (with_large) {
                for (int j = l_off[b]; j < l_off[b + 1]; ++j) { sh.l_pos.push_back(map_row(l_pos_all[j])); sh.l_z.push_back(l_z_all[j]); }
                sh.l_off.push_back((int32_t)sh.l_pos.size());
            }
        }
        sh.beta_s.assign(sh.s_pos.size() * n_folds + 1, 0.0);
        sh.beta_l.assign(sh.l_pos.size() * n_folds + 1, 0.0);
        sh.status.assign(sh.blocks.size() + 1, 0);
    }

    // ---- variance side channel inputs (dbslmm.cpp:265-275, 300-304, 319-320): test .bim positions, indicator, test .bed
    vector<int32_t> test_indicator;
    vector<uint8_t> test_bed;
    int64_t test_n_snp = 0;
    int n_test = 0;
    if (want_var) {
        ifstream ind(cPar.test_indicator_file.c_str());
        if (!ind) { cerr << "ERROR: " << cPar.test_indicator_file << " dose not exist!" << endl; exit(1); }
        string line;
        while (getline(ind, line)) test_indicator.push_back(atoi(line.c_str()));                 // read_indices_file: first field
        for (int v : test_indicator) n_test += (v != 0);
        unordered_map<long, int32_t> ps2row;                                                      // readTestBim + makePosObjectForTestBim
        ifstream tb((cPar.dat_str + ".bim").c_str());
        if (!tb) { cerr << "ERROR: " << cPar.dat_str << ".bim dose not exist!" << endl; exit(1); }
        while (getline(tb, line)) {
            vector<string> t; stringstream ss(line); string e;
            while (getline(ss, e, '\t')) t.push_back(e);
            if (t.size() >= 4) ps2row.emplace(atol(t[3].c_str()), (int32_t)test_n_snp);          // first match wins, like the linear scan
            test_n_snp++;
        }
        if (!read_bed(cPar.dat_str + ".bed", test_n_snp, (int)test_indicator.size(), test_bed)) {
            cerr << "ERROR: cannot read SNP-major " << cPar.dat_str << ".bed" << endl; exit(1);
        }
        auto lookup = [&](long ps, const string& snp) -> int32_t {
            auto it = ps2row.find(ps);
            if (it == ps2row.end()) { cerr << "ERROR: SNP " << snp << " (position " << ps << ") is not in " << cPar.dat_str << ".bim" << endl; exit(1); }
            return it->second;
        };
        const Job& j = jobs[0];
        vector<int32_t> s_tp(j.info_s.size()), l_tp(j.info_l.size());
        for (size_t i = 0; i < j.info_s.size(); ++i) s_tp[i] = lookup(j.info_s.ps[i], j.info_s.snp[i]);
        for (size_t i = 0; i < j.info_l.size(); ++i) l_tp[i] = lookup(j.info_l.ps[i], j.info_l.snp[i]);
        for (int g = 0; g < n_gpus; ++g) {
            Shard& sh = shards[g];
            for (int b : sh.blocks) {
                for (int q = s_off[b]; q < s_off[b + 1]; ++q) sh.s_tpos.push_back(s_tp[q]);
                if (with_large) for (int q = l_off[b]; q < l_off[b + 1]; ++q) sh.l_tpos.push_back(l_tp[q]);
            }
            sh.variance.assign((size_t)n_folds * sh.blocks.size() * std::max(n_test, 1) + 1, 0.0);
        }
    }

    // ---- fit (the reference's "Fitting time" window, dbslmm.cpp:331-350 / 372-388)
    const double t_fitting = walltime();
    cout << "Fitting model..." << endl;
    auto run = [&](int g) {
        Shard& sh = shards[g];
        if (sh.blocks.empty()) return;
        dbslmm_b200_fit_args a{};
        // the panel (this GPU's shard of it) travels with the fit, like est()'s bed_str: uploaded in batches that
        // overlap decode / Gram / Cholesky.  One GPU with a MAF constraint already holds the panel (pre-pass above).
        if (n_gpus > 1) { a.bed = bed; a.bed_n_snp = n_snp_all; a.bed_n_ref = n_ref; }
        else if (!early_load) { a.bed = bed; a.bed_n_snp = n_snp_all; a.bed_n_ref = n_ref; }
        a.n_blocks = (int32_t)sh.blocks.size();
        a.s_off = sh.s_off.data(); a.s_pos = sh.s_pos.data(); a.s_z = sh.s_z.data();
        if (with_large) { a.l_off = sh.l_off.data(); a.l_pos = sh.l_pos.data(); a.l_z = sh.l_z.data(); }
        a.n_folds = n_folds; a.sigma_s = sigma_s.data(); a.n_obs = cPar.n; a.tau = cPar.tau;
        a.solver = solver; a.flags = (n_gpus > 1) ? DBSLMM_B200_FLAG_PANEL_SUBSET : 0;
        a.beta_s_out = sh.beta_s.data(); a.beta_l_out = with_large ? sh.beta_l.data() : nullptr;
        a.block_status_out = sh.status.data(); a.timing = &sh.timing;
        if (want_var) {
            a.test_bed = test_bed.data(); a.test_n_snp = test_n_snp; a.test_n_total = (int32_t)test_indicator.size();
            a.test_indicator = test_indicator.data(); a.s_tpos = sh.s_tpos.data();
            a.l_tpos = with_large ? sh.l_tpos.data() : nullptr; a.variance_out = sh.variance.data();
        }
        sh.rc = dbslmm_b200_fit(hs[g], &a);
        if (sh.rc < 0) sh.err = dbslmm_b200_last_error(hs[g]);
    };
    // Several GPUs without the variance side channel: ONE library call fans the blocks out (dbslmm_b200_fit_multi) and
    // returns block-major betas; the per-GPU shards above serve the variance side channel and the single-GPU run.
    const bool use_multi = n_gpus > 1 && !want_var;
    vector<double> multi_s, multi_l;
    if (use_multi) {
        multi_s.assign(s_pos_all.size() * n_folds + 1, 0.0);
        multi_l.assign(l_pos_all.size() * n_folds + 1, 0.0);
        dbslmm_b200_fit_args a{};
        a.bed = bed; a.bed_n_snp = n_snp_all; a.bed_n_ref = n_ref;
        a.n_blocks = num_block;
        a.s_off = s_off.data(); a.s_pos = s_pos_all.data(); a.s_z = s_z_all.data();
        if (with_large) { a.l_off = l_off.data(); a.l_pos = l_pos_all.data(); a.l_z = l_z_all.data(); }
        a.n_folds = n_folds; a.sigma_s = sigma_s.data(); a.n_obs = cPar.n; a.tau = cPar.tau;
        a.solver = solver;
        a.beta_s_out = multi_s.data(); a.beta_l_out = with_large ? multi_l.data() : nullptr;
        a.timing = &shards[0].timing;
        const int rc = dbslmm_b200_fit_multi(hs.data(), n_gpus, &a);
        if (rc < 0) { cerr << "ERROR: " << dbslmm_b200_last_error(hs[0]) << endl; exit(2); }
        if (rc > 0) cerr << "ERROR: Matrix is Singular! (" << rc << " block(s))" << endl;                  // dbslmmfit.cpp:665
    } else {
        vector<thread> th;
        for (int g = 1; g < n_gpus; ++g) th.emplace_back(run, g);
        run(0);
        for (auto& t : th) t.join();
        for (int g = 0; g < n_gpus; ++g) {
            if (shards[g].rc < 0) { cerr << "ERROR: GPU " << g << ": " << shards[g].err << endl; exit(2); }
            if (shards[g].rc > 0) cerr << "ERROR: Matrix is Singular! (" << shards[g].rc << " block(s) on GPU " << g << ")" << endl;   // dbslmmfit.cpp:665
        }
    }
    const double time_fitting = walltime() - t_fitting;
    cout << "Fitting time: " << time_fitting << " seconds." << endl;
    if (cPar.verbose)
        for (int g = 0; g < n_gpus; ++g) {
            const dbslmm_b200_timing& t = shards[g].timing;
            cout << "[gpu " << g << "] blocks " << shards[g].blocks.size() << " h2d " << t.h2d_ms << " decode " << t.decode_ms << " gram "
                 << t.gram_ms << " solve " << t.solve_ms << " d2h " << t.d2h_ms << " ms, " << t.n_launches << " launches" << endl;
        }

    // ---- gather to block-major global order
    const size_t tot_s = s_pos_all.size(), tot_l = l_pos_all.size();
    vector<double> beta_s(tot_s * n_folds + 1), beta_l(tot_l * n_folds + 1);
    if (use_multi) { beta_s.swap(multi_s); beta_l.swap(multi_l); }
    for (int g = 0; g < n_gpus && !use_multi; ++g) {
        const Shard& sh = shards[g];
        const size_t ns = sh.s_pos.size(), nl = sh.l_pos.size();
        for (size_t i = 0; i < sh.blocks.size(); ++i) {
            const int b = sh.blocks[i];
            for (int f = 0; f < n_folds; ++f) {
                for (int j = 0; j < m_s[b]; ++j) beta_s[f * tot_s + s_off[b] + j] = sh.beta_s[f * ns + sh.s_off[i] + j];
                if (with_large)
                    for (int j = 0; j < m_l[b]; ++j) beta_l[f * tot_l + l_off[b] + j] = sh.beta_l[f * nl + sh.l_off[i] + j];
            }
        }
    }

    // ---- output effect per chromosome (dbslmm.cpp:353-364 / 391-395): large first (flag 1), then small (flag 0);
    // the chromosomes' files are formatted and written in parallel ("%g" = the default ostream formatting of a double)
    const double t_out0 = walltime();
    parallel_for((int)jobs.size(), host_threads, [&](int ji) {
        const Job& j = jobs[ji];
        const size_t js = (size_t)s_off[j.block0], jl = (size_t)l_off[j.block0];   // this chromosome's slice of the global arrays
        for (int f = 0; f < n_folds; ++f) {
            string eff_str = j.eff + ".txt";
            if (n_folds > 1) { ostringstream o; o << j.eff << "_f" << f << ".txt"; eff_str = o.str(); }
            string txt;
            txt.reserve((j.info_s.size() + j.info_l.size()) * 48 + 64);
            char num[80];
            auto row = [&](const string& snp, const string& a1, double bta, double noscl, int flag) {
                txt += snp; txt += ' '; txt += a1; txt += ' ';
                const int n = snprintf(num, sizeof num, "%g %g %d\n", bta, noscl, flag);
                txt.append(num, (size_t)n);
            };
            if (j.with_large)
                for (size_t i = 0; i < j.info_l.size(); ++i) {
                    const double bta = beta_l[f * tot_l + jl + i];
                    const double noscl = bta / sqrt(2 * j.info_l.maf[i] * (1 - j.info_l.maf[i]));
                    if (isinf(noscl) == false) row(j.info_l.snp[i], j.info_l.a1[i], bta, noscl, 1);
                }
            for (size_t i = 0; i < j.info_s.size(); ++i) {
                const double bta = beta_s[f * tot_s + js + i];
                const double noscl = bta / sqrt(2 * j.info_s.maf[i] * (1 - j.info_s.maf[i]));
                if (j.info_s.snp[i].size() != 0 && isinf(noscl) == false) row(j.info_s.snp[i], j.info_s.a1[i], bta, noscl, 0);
            }
            ofstream effFout(eff_str.c_str(), ios::binary);
            effFout.write(txt.data(), (streamsize)txt.size());
            effFout.close();
        }
    });
    const double t_out1 = walltime();
    if (cPar.verbose || jobs.size() > 1)
        cout << "[timing] panel files " << (t_panel - t_start) << " s, summary statistics + matching " << (t_ingest - t_panel)
             << " s, plan + fit " << (t_out0 - t_ingest) << " s (Fitting time " << time_fitting << " s), output " << (t_out1 - t_out0)
             << " s; " << host_threads << " host thread(s), " << n_gpus << " GPU(s)" << endl;
    if (want_var) {
        // diags.save("variance.txt", arma_ascii) (dbslmmfit.cpp:242, 361): n_test x num_block, cwd-relative
        for (int f = 0; f < n_folds; ++f) {
            vector<double> diags((size_t)n_test * num_block, 0.0);
            for (int g = 0; g < n_gpus; ++g) {
                const Shard& sh = shards[g];
                for (size_t i = 0; i < sh.blocks.size(); ++i)
                    for (int t = 0; t < n_test; ++t)
                        diags[(size_t)t * num_block + sh.blocks[i]] = sh.variance[((size_t)f * sh.blocks.size() + i) * n_test + t];
            }
            string name = "variance.txt";
            if (n_folds > 1) { ostringstream o; o << "variance_f" << f << ".txt"; name = o.str(); }
            ofstream vf(name.c_str());
            vf << "ARMA_MAT_TXT_FN008" << '\n' << n_test << ' ' << num_block << '\n';
            vf.setf(ios::scientific);
            vf.precision(16);
            for (int t = 0; t < n_test; ++t) {
                for (int b = 0; b < num_block; ++b) { vf.put(' '); vf.width(24); vf << diags[(size_t)t * num_block + b]; }
                vf.put('\n');
            }
        }
    }
    if (!cPar.dump_bin.empty()) {
        // int64 n_folds, tot_l, tot_s, then FP64 beta_l[n_folds][tot_l], beta_s[n_folds][tot_s] (chromosomes in manifest order)
        ofstream o(cPar.dump_bin.c_str(), ios::binary);
        const int64_t hdr[3] = {n_folds, (int64_t)tot_l, (int64_t)tot_s};
        o.write((const char*)hdr, sizeof(hdr));
        o.write((const char*)beta_l.data(), sizeof(double) * tot_l * n_folds);
        o.write((const char*)beta_s.data(), sizeof(double) * tot_s * n_folds);
    }
    dbslmm_b200_host_free(hs[0], bed);
    for (auto* h : hs) dbslmm_b200_destroy(h);
    return EXIT_SUCCESS;
}
