// ingest.hpp -- host-side ingest of the `dbslmm` command line: the text readers and matchers
// that decide WHICH SNPs and blocks reach the GPU.  Behaviour-compatible restatement of the
// reference's IO / SNPPROC classes (scr/dtpr.cpp:47-68, 71-80, 83-123, 178-220, 383-408,
// 455-481) with SoA outputs instead of vectors of string-laden structs.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

namespace dbslmm_host {

struct BimEntry { int64_t pos; std::string a1, a2; };
using BimMap = std::unordered_map<std::string, BimEntry>;

struct Block { long start, end; };

struct Summ {                       // SUMM, dtpr.hpp:58-68 (P is never used by the fit and is not computed)
    std::vector<std::string> snp, a1, a2;
    std::vector<long> ps;
    std::vector<double> maf, z;
    size_t size() const { return snp.size(); }
};

struct Info {                       // INFO (dtpr.hpp:115-125) as SoA, block-sorted
    std::vector<std::string> snp, a1;
    std::vector<long> ps;
    std::vector<int32_t> pos, block;
    std::vector<double> maf, z;
    size_t size() const { return snp.size(); }
};

int get_row(const std::string& path);                                       // IO::getRow
bool read_block(const std::string& path, std::vector<Block>& out);          // IO::readBlock
int64_t read_bim(const std::string& path, BimMap& out);                     // IO::readBim (map part)
bool read_summ(const std::string& path, Summ& out);                         // IO::readSumm
bool read_bed(const std::string& path, int64_t n_snp, int n_ref, std::vector<uint8_t>& out);
// the same into caller memory (n_snp * ceil(n_ref / 4) bytes), e.g. a slice of one pinned genome-wide buffer
bool read_bed_into(const std::string& path, int64_t n_snp, int n_ref, uint8_t* dst);

// SNPPROC::matchRef: alleles must match exactly, |maf_ref - maf_summ| < mafMax (ref_maf == nullptr
// when the MAF pre-pass is disabled: the reference then compares against 0).  `matched[i]` mirrors
// badsnp_bool.  Returns dis/maf discrepancy counts through the out-params, like the reference's log.
void match_ref(const Summ& summ, const BimMap& bim, const double* ref_maf, double maf_max, Info& inter,
               std::vector<char>& matched, int& dis_count, int& maf_count);

// SNPPROC::addBlock: start <= ps < end with the sorted-input early break; SNPs the scan never
// reaches are dropped (the reference leaves them as empty records that its writer skips).
int add_block(const Info& inter, const std::vector<Block>& blocks, Info& out);

// count_snps_per_block (helpers.cpp:16-30) as CSR offsets
std::vector<int32_t> block_offsets(const Info& info, int n_blocks);

}  // namespace dbslmm_host
