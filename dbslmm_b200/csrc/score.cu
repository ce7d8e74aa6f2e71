// score.cu -- polygenic-score accumulation over a validation panel (BASELINE config 4).
//
// Replaces the `plink --bfile val --score <eff>.txt 1 2 4 sum` step that the reference's driver runs
// once per heritability fold (DBSLMM_script.sh:87, scored column 4 = beta / sqrt(2 maf (1-maf)) written
// by scr/dbslmm.cpp:354-362): score_i = sum_j beta_j * dosage_ij, dosage = copies of the scored
// allele, missing calls mean-imputed (PLINK's default), all folds in ONE pass over the 2-bit .bed.
// PLINK is external to the reference, so this path's parity is pinned only against a numpy
// restatement of those documented semantics (tests/test_gpu_score.py).
//
// Every .bed byte is read once (coalesced, 4 bytes = 16 individuals per thread); per-chunk partial sums are written to
// a scratch buffer and reduced in a fixed order (deterministic).  The work is one FP64 multiply-add per genotype and
// fold (3.3e10 at config 4), so the FP64 pipe, not HBM, bounds it.
#include "common.cuh"
#include "kernels.h"

namespace dbslmm {

static constexpr int kScoreThreads = 256;
static constexpr int kMaxFolds = 3;          // folds per pass (more folds => several passes): 16 individuals x 3 folds of FP64 accumulators per thread

// One thread = FOUR consecutive .bed bytes = 16 individuals (a warp reads 128 contiguous bytes of a row per load), 64
// scored SNPs per shared-memory batch.  Per SNP the four possible dosages {code 0, 1 = missing -> mean, 2, 3} are put
// in shared memory once (allele flip and mean imputation folded in), so the inner loop is: extract the 2-bit code,
// load the dosage it selects (one LDS.64), NF fused multiply-adds.
template <int NF>
__global__ void __launch_bounds__(kScoreThreads, 2)
prs_partial_kernel(const uint8_t* __restrict__ bed, int32_t pitch, int32_t n_val, const SnpStat* __restrict__ stats,
                   const int32_t* __restrict__ pos, const uint8_t* __restrict__ flip, const double* __restrict__ beta,
                   int64_t beta_stride, int32_t n_scored, int32_t rows_per_chunk, double* __restrict__ partial) {
    __shared__ double s_beta[64][NF];
    __shared__ __align__(16) double s_dos[64][4];            // dosage of code 0, 1, 2, 3
    __shared__ int32_t s_row[64];
    const int tid = threadIdx.x;
    const int word_col = blockIdx.x * kScoreThreads + tid;   // 4-byte column of the row
    const int byte0 = word_col * 4;
    const bool in_range = byte0 < pitch;
    const bool full_word = byte0 + 4 <= pitch && (pitch & 3) == 0;     // rows stay 4-byte aligned only if the pitch is
    const int s_begin = blockIdx.y * rows_per_chunk;
    const int s_end = min(n_scored, s_begin + rows_per_chunk);
    double acc[16][NF];
#pragma unroll
    for (int q = 0; q < 16; ++q)
#pragma unroll
        for (int f = 0; f < NF; ++f) acc[q][f] = 0.0;
    for (int s0 = s_begin; s0 < s_end; s0 += 64) {
        const int nb = min(64, s_end - s0);
        __syncthreads();
        if (tid < nb) {
            const int row = pos[s0 + tid];
            const SnpStat st = stats[row];
            const bool fl = flip != nullptr && flip[s0 + tid] != 0;
            const double mu = (double)st.sum / (double)st.n_nonmiss;        // mean A1 dosage of called genotypes
            s_row[tid] = row;
            // A1 copies: code 0 -> 2, 2 -> 1, 3 -> 0, 1 -> missing (dtpr.cpp:329-350)
            s_dos[tid][0] = fl ? 0.0 : 2.0;
            s_dos[tid][1] = fl ? 2.0 - mu : mu;
            s_dos[tid][2] = 1.0;
            s_dos[tid][3] = fl ? 2.0 : 0.0;
#pragma unroll
            for (int f = 0; f < NF; ++f) s_beta[tid][f] = beta[(int64_t)f * beta_stride + s0 + tid];
        }
        __syncthreads();
        if (!in_range) continue;
        // software pipeline: the four row words of the NEXT group are in flight while this group is accumulated
        auto load_word = [&](int j) -> uint32_t {
            if (j >= nb) return 0xFFFFFFFFu;
            const uint8_t* p = bed + (size_t)s_row[j] * pitch + byte0;
            if (full_word) return __ldg(reinterpret_cast<const uint32_t*>(p));
            uint32_t w = 0;
            for (int bq = 0; bq < 4; ++bq) w |= (uint32_t)((byte0 + bq < pitch) ? p[bq] : 0xFF) << (8 * bq);
            return w;
        };
        uint32_t words[4], next[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) next[u] = load_word(u);
        for (int j0 = 0; j0 < nb; j0 += 4) {
#pragma unroll
            for (int u = 0; u < 4; ++u) words[u] = next[u];
#pragma unroll
            for (int u = 0; u < 4; ++u) next[u] = load_word(j0 + 4 + u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (j0 + u >= nb) break;
                const uint32_t w = words[u];
                // the SNP's four dosages sit in 32 contiguous bytes of shared memory: one broadcast-class LDS.64 indexed
                // by the 2-bit code replaces the compare/select chain
                const char* dos = reinterpret_cast<const char*>(&s_dos[j0 + u][0]);
                double bf[NF];
#pragma unroll
                for (int f = 0; f < NF; ++f) bf[f] = s_beta[j0 + u][f];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const uint32_t c8 = (q < 15) ? ((w >> (2 * q)) & 3u) << 3 : (w >> 30) << 3;     // code * 8 bytes
                    const double d = *reinterpret_cast<const double*>(dos + c8);
#pragma unroll
                    for (int f = 0; f < NF; ++f) acc[q][f] = fma(bf[f], d, acc[q][f]);
                }
            }
        }
    }
    if (!in_range) return;
#pragma unroll
    for (int f = 0; f < NF; ++f)
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int ind = byte0 * 4 + q;
            if (ind < n_val) partial[((size_t)blockIdx.y * NF + f) * n_val + ind] = acc[q][f];
        }
}

__global__ void prs_reduce_kernel(const double* __restrict__ partial, int32_t n_chunks, int32_t nf, int32_t n_val,
                                  double* __restrict__ scores) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (i >= n_val) return;
    double s = 0.0;
    for (int c = 0; c < n_chunks; ++c) s += partial[((size_t)c * nf + f) * n_val + i];
    scores[(size_t)f * n_val + i] = s;
}

cudaError_t launch_prs(const uint8_t* bed, int32_t n_val, const SnpStat* stats, const int32_t* pos, const uint8_t* flip,
                       const double* beta, int64_t beta_stride, int32_t n_scored, int32_t nf, int32_t n_chunks,
                       double* partial, double* scores, cudaStream_t st) {
    if (n_scored == 0 || nf == 0) return cudaSuccess;
    const int32_t pitch = (n_val + 3) / 4;
    const int rows_per_chunk = (n_scored + n_chunks - 1) / n_chunks;
    dim3 grid(((pitch + 3) / 4 + kScoreThreads - 1) / kScoreThreads, n_chunks);
    switch (nf) {
        case 1: prs_partial_kernel<1><<<grid, kScoreThreads, 0, st>>>(bed, pitch, n_val, stats, pos, flip, beta, beta_stride, n_scored, rows_per_chunk, partial); break;
        case 2: prs_partial_kernel<2><<<grid, kScoreThreads, 0, st>>>(bed, pitch, n_val, stats, pos, flip, beta, beta_stride, n_scored, rows_per_chunk, partial); break;
        case 3: prs_partial_kernel<3><<<grid, kScoreThreads, 0, st>>>(bed, pitch, n_val, stats, pos, flip, beta, beta_stride, n_scored, rows_per_chunk, partial); break;
        default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 rgrid((n_val + 255) / 256, nf);
    prs_reduce_kernel<<<rgrid, 256, 0, st>>>(partial, n_chunks, nf, n_val, scores);
    return cudaGetLastError();
}

}  // namespace dbslmm
