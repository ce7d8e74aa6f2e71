// ingest.cpp -- see ingest.hpp.  Quirks kept on purpose: no header skipping,
// no allele flipping, first duplicate rs id wins, atof/atoi/atol prefix parsing, tab separator.
#include "ingest.hpp"

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace dbslmm_host {

static void split_tab(const std::string& line, std::vector<std::string>& out) {
    out.clear();
    std::stringstream ss(line);
    std::string el;
    while (std::getline(ss, el, '\t')) out.push_back(el);
}

int get_row(const std::string& path) {                       // dtpr.cpp:71-80
    std::ifstream f(path.c_str());
    std::string row;
    int n = 0;
    while (std::getline(f, row)) n++;
    return n;
}

bool read_block(const std::string& path, std::vector<Block>& out) {   // dtpr.cpp:47-68
    std::ifstream f(path.c_str());
    if (!f) return false;
    std::string line;
    std::vector<std::string> t;
    while (std::getline(f, line)) {
        split_tab(line, t);
        if (t.size() < 3) continue;
        out.push_back({atol(t[1].c_str()), atol(t[2].c_str())});
    }
    return true;
}

int64_t read_bim(const std::string& path, BimMap& out) {     // dtpr.cpp:107-121
    std::ifstream f(path.c_str());
    std::string line;
    std::vector<std::string> t;
    int64_t count = 0;
    while (std::getline(f, line)) {
        split_tab(line, t);
        if (t.size() >= 6) out.emplace(t[1], BimEntry{count, t[4], t[5]});   // emplace keeps the first duplicate
        count++;
    }
    return count;
}

bool read_summ(const std::string& path, Summ& out) {         // dtpr.cpp:178-220
    std::ifstream f(path.c_str());
    if (!f) return false;
    std::string line;
    std::vector<std::string> t;
    while (std::getline(f, line)) {
        split_tab(line, t);
        if (t.size() < 11) continue;                         // the reference would index out of range here
        double z = 0.0;
        if (isdigit((unsigned char)t[9].c_str()[0])) {       // :194
            const double se = atof(t[9].c_str());
            if (se - 0.0 > 1e-20) z = atof(t[8].c_str()) / se;   // :196-197
        }
        out.snp.push_back(t[1]);
        out.ps.push_back(atol(t[2].c_str()));
        out.a1.push_back(t[5]);
        out.a2.push_back(t[6]);
        const double af = atof(t[7].c_str());
        out.maf.push_back(std::min(af, 1.0 - af));           // :213
        out.z.push_back(z);
    }
    return true;
}

bool read_bed(const std::string& path, int64_t n_snp, int n_ref, std::vector<uint8_t>& out) {
    std::ifstream f(path.c_str(), std::ios::binary);
    if (!f) return false;
    unsigned char magic[3] = {0, 0, 0};
    f.read(reinterpret_cast<char*>(magic), 3);
    if (magic[0] != 0x6C || magic[1] != 0x1B || magic[2] != 0x01) return false;   // SNP-major only
    const size_t bytes = (size_t)n_snp * (size_t)((n_ref + 3) / 4);
    out.resize(bytes);
    f.read(reinterpret_cast<char*>(out.data()), (std::streamsize)bytes);
    return (size_t)f.gcount() == bytes;
}

void match_ref(const Summ& summ, const BimMap& bim, const double* ref_maf, double maf_max, Info& inter,
               std::vector<char>& matched, int& dis_count, int& maf_count) {       // dtpr.cpp:383-408
    dis_count = maf_count = 0;
    matched.assign(summ.size(), 0);
    for (size_t i = 0; i < summ.size(); ++i) {
        auto it = bim.find(summ.snp[i]);
        // an unknown rs id default-constructs an ALLELE in the reference (empty alleles, maf 0)
        const bool known = it != bim.end();
        const bool a1_ok = known ? it->second.a1 == summ.a1[i] : summ.a1[i].empty();
        const bool a2_ok = known ? it->second.a2 == summ.a2[i] : summ.a2[i].empty();
        const double rmaf = (known && ref_maf) ? ref_maf[it->second.pos] : 0.0;
        const bool maf_ok = std::fabs(rmaf - summ.maf[i]) < maf_max;
        if (!a1_ok || !a2_ok) dis_count++;
        if (!maf_ok) maf_count++;
        if (known && a1_ok && a2_ok && maf_ok) {
            inter.snp.push_back(summ.snp[i]);
            inter.a1.push_back(summ.a1[i]);
            inter.ps.push_back(summ.ps[i]);
            inter.pos.push_back((int32_t)it->second.pos);
            inter.block.push_back(-1);
            inter.maf.push_back(summ.maf[i]);
            inter.z.push_back(summ.z[i]);
            matched[i] = 1;
        }
    }
}

int add_block(const Info& inter, const std::vector<Block>& blocks, Info& out) {   // dtpr.cpp:455-481
    const int nb = (int)blocks.size();
    size_t count = 0;
    for (int i = 0; i < nb; ++i) {
        const int start = (int)blocks[i].start, end = (int)blocks[i].end;         // the reference narrows to int
        for (size_t j = count; j < inter.size(); ++j) {
            if (inter.ps[j] >= start && inter.ps[j] < end) {
                out.snp.push_back(inter.snp[j]);
                out.a1.push_back(inter.a1[j]);
                out.ps.push_back(inter.ps[j]);
                out.pos.push_back(inter.pos[j]);
                out.block.push_back(i);
                out.maf.push_back(inter.maf[j]);
                out.z.push_back(inter.z[j]);
                count++;
            } else {
                break;
            }
        }
    }
    return nb;
}

std::vector<int32_t> block_offsets(const Info& info, int n_blocks) {              // helpers.cpp:16-30
    std::vector<int32_t> off((size_t)n_blocks + 1, 0);
    for (size_t j = 0; j < info.size(); ++j) off[(size_t)info.block[j] + 1]++;
    for (int b = 0; b < n_blocks; ++b) off[b + 1] += off[b];
    return off;
}

}  // namespace dbslmm_host
