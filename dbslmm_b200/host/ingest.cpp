// ingest.cpp -- see ingest.hpp.  Quirks kept on purpose: no header skipping,
// no allele flipping, first duplicate rs id wins, atof/atoi/atol prefix parsing, tab separator.
#include "ingest.hpp"

#include <algorithm>
#include <cctype>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

namespace dbslmm_host {

// ---- fast text readers -------------------------------------------------------------------------------------
// The reference reads its inputs with getline + stringstream per line (dtpr.cpp:47-220); at genome scale (1.1 M-line
// files) that costs seconds.  Here a file is read in one go and walked with pointers; the SEMANTICS of the
// reference's tokenising and number parsing are kept:
//   * lines end at '\n' (a '\r' stays in the last field); a last line without newline counts;
//   * fields are split at tabs the way getline(ss, el, '\t') does it: empty fields in the middle are kept, an empty
//     field after a trailing tab is not produced, an empty line has no fields;
//   * atol / atoi / atof prefix parsing: leading white space, optional sign, the longest valid prefix, 0 if none.
struct Tok { const char* b; const char* e; size_t size() const { return (size_t)(e - b); } };

static bool slurp(const std::string& path, std::string& out) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::fseek(f, 0, SEEK_END);
    const long n = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(n > 0 ? (size_t)n : 0);
    size_t got = 0;
    if (n > 0) got = std::fread(&out[0], 1, (size_t)n, f);
    std::fclose(f);
    out.resize(got);
    return true;
}

// next line of [p, end): returns false at the end of the buffer
static inline bool next_line(const char*& p, const char* end, Tok& line) {
    if (p >= end) return false;
    const char* nl = (const char*)std::memchr(p, '\n', (size_t)(end - p));
    line.b = p;
    line.e = nl ? nl : end;
    p = nl ? nl + 1 : end;
    return true;
}

static inline void split_tab(const Tok& line, std::vector<Tok>& out) {
    out.clear();
    const char* p = line.b;
    while (p < line.e) {
        const char* t = (const char*)std::memchr(p, '\t', (size_t)(line.e - p));
        if (!t) { out.push_back({p, line.e}); return; }
        out.push_back({p, t});
        p = t + 1;
    }
}

static inline long tok_atol(const Tok& t) {
    const char* p = t.b;
    while (p < t.e && std::isspace((unsigned char)*p)) ++p;
    bool neg = false;
    if (p < t.e && (*p == '+' || *p == '-')) { neg = (*p == '-'); ++p; }
    unsigned long v = 0;
    while (p < t.e && *p >= '0' && *p <= '9') { v = v * 10 + (unsigned long)(*p - '0'); ++p; }
    return neg ? -(long)v : (long)v;
}

static inline double tok_atof(const Tok& t) {
    // plain decimal numbers ([-]digits[.digits][e[+-]digits], the whole field) go through from_chars (correctly rounded,
    // like strtod); anything else -- leading blanks, '+', hex, inf/nan, trailing junk -- takes atof on a copy
    const char* p = t.b;
    if (p < t.e && *p != '+' && !std::isspace((unsigned char)*p)) {
        double v = 0.0;
        const std::from_chars_result r = std::from_chars(t.b, t.e, v, std::chars_format::general);
        if (r.ec == std::errc() && r.ptr == t.e) {
            // from_chars also accepts "inf"/"nan" spellings that atof accepts too; both give the same value
            return v;
        }
    }
    char buf[64];
    const size_t n = std::min(t.size(), sizeof(buf) - 1);
    std::memcpy(buf, t.b, n);
    buf[n] = 0;
    if (t.size() < sizeof(buf)) return std::atof(buf);
    return std::atof(std::string(t.b, t.e).c_str());
}

int get_row(const std::string& path) {                       // dtpr.cpp:71-80
    std::string buf;
    if (!slurp(path, buf)) return 0;
    const char* p = buf.data();
    const char* end = p + buf.size();
    Tok line;
    int n = 0;
    while (next_line(p, end, line)) n++;
    return n;
}

bool read_block(const std::string& path, std::vector<Block>& out) {   // dtpr.cpp:47-68
    std::string buf;
    if (!slurp(path, buf)) return false;
    const char* p = buf.data();
    const char* end = p + buf.size();
    Tok line;
    std::vector<Tok> t;
    while (next_line(p, end, line)) {
        split_tab(line, t);
        if (t.size() < 3) continue;
        out.push_back({tok_atol(t[1]), tok_atol(t[2])});
    }
    return true;
}

int64_t read_bim(const std::string& path, BimMap& out) {     // dtpr.cpp:107-121
    std::string buf;
    if (!slurp(path, buf)) return 0;
    const char* p = buf.data();
    const char* end = p + buf.size();
    out.reserve((size_t)(buf.size() / 24));
    Tok line;
    std::vector<Tok> t;
    int64_t count = 0;
    while (next_line(p, end, line)) {
        split_tab(line, t);
        if (t.size() >= 6)      // emplace keeps the first duplicate
            out.emplace(std::string(t[1].b, t[1].e), BimEntry{count, std::string(t[4].b, t[4].e), std::string(t[5].b, t[5].e)});
        count++;
    }
    return count;
}

bool read_summ(const std::string& path, Summ& out) {         // dtpr.cpp:178-220
    std::string buf;
    if (!slurp(path, buf)) return false;
    const char* p = buf.data();
    const char* end = p + buf.size();
    const size_t guess = buf.size() / 60 + 16;
    out.snp.reserve(guess); out.a1.reserve(guess); out.a2.reserve(guess);
    out.ps.reserve(guess); out.maf.reserve(guess); out.z.reserve(guess);
    Tok line;
    std::vector<Tok> t;
    while (next_line(p, end, line)) {
        split_tab(line, t);
        if (t.size() < 11) continue;                         // the reference would index out of range here
        double z = 0.0;
        if (t[9].size() > 0 && isdigit((unsigned char)t[9].b[0])) {      // :194
            const double se = tok_atof(t[9]);
            if (se - 0.0 > 1e-20) z = tok_atof(t[8]) / se;   // :196-197
        }
        out.snp.emplace_back(t[1].b, t[1].e);
        out.ps.push_back(tok_atol(t[2]));
        out.a1.emplace_back(t[5].b, t[5].e);
        out.a2.emplace_back(t[6].b, t[6].e);
        const double af = tok_atof(t[7]);
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

bool read_bed_into(const std::string& path, int64_t n_snp, int n_ref, uint8_t* dst) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    unsigned char magic[3] = {0, 0, 0};
    bool ok = std::fread(magic, 1, 3, f) == 3 && magic[0] == 0x6C && magic[1] == 0x1B && magic[2] == 0x01;   // SNP-major only
    const size_t bytes = (size_t)n_snp * (size_t)((n_ref + 3) / 4);
    if (ok) ok = std::fread(dst, 1, bytes, f) == bytes;
    std::fclose(f);
    return ok;
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
