// variance.cu -- the fork's asymptotic-variance side channel without a single dense m x m GEMM.
//
// Reference: calc_nt_by_nt_matrix (scr/calc_asymptotic_variance.cpp:22-57) forms Ainv = (Sigma_ss + cI)^-1,
// var_bl = S^-1 / n, var_bs = n sigma^4 (mat1 + mat2 n var_bl mat2') with four to six m_s^3 products and then
// X_l var_bl X_l' + X_s var_bs X_s' (n_test x n_test), of which only the diagonal is used
// (scr/dbslmmfit.cpp:538,625).  With A = Sigma_ss + cI one has mat1 = c (I - c Ainv), mat2 = c Ainv Sigma_sl, and
// in terms of the Cholesky factor L of the bordered matrix K (chol.cu):
//     d_i = |L22^-1 x_l,i|^2 / n + sigma^2 ( |x_s,i|^2 - c |L11^-1 x_s,i|^2 + c |L22^-1 L21 L11^-1 x_s,i|^2 )
// Both forward substitutions come for free from the left-looking factorisation: the standardised test genotypes
// are appended as extra matrix rows  A_i = [x_s,i | 0]  and  B_i = [0 | x_l,i]  below the z row, so after the panel
// steps those rows hold  [L11^-1 x_s ; -L22^-1 L21 L11^-1 x_s]  and  [0 ; L22^-1 x_l].  This file only fills the rows
// and takes the row norms.  Test genotypes are standardised WITHIN the selected test subset with mean imputation
// (scr/dbslmmfit.cpp:427-429 -> dtpr.cpp:285-380).
#include "common.cuh"
#include "kernels.h"

namespace dbslmm {

// per SNP row g (block order): mean and 1/sd of the selected test individuals
__global__ void test_stats_kernel(const uint8_t* __restrict__ tbed, int32_t tpitch, const int32_t* __restrict__ sel,
                                  int32_t n_test, const int32_t* __restrict__ tpos, int64_t n_rows,
                                  double* __restrict__ tmu, double* __restrict__ tisd) {
    const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n_rows) return;
    const int lane = threadIdx.x & 31;
    const uint8_t* row = tbed + (size_t)tpos[g] * tpitch;
    int cnt = 0, sum = 0, sq = 0;
    for (int i = lane; i < n_test; i += 32) {
        const int s = sel[i];
        const unsigned c = (row[s >> 2] >> (2 * (s & 3))) & 3u;
        const int d = (c == 0u) ? 2 : (c == 2u) ? 1 : 0;
        if (c != 1u) { cnt++; sum += d; sq += d * d; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
    }
    if (lane == 0) {
        const double mu = (double)sum / (double)cnt;
        const double var = ((double)sq - (double)cnt * mu * mu) / (double)(n_test - 1);   // imputed entries add 0
        tmu[g] = mu;
        tisd[g] = 1.0 / sqrt(var);
    }
}

// rows A_i (and B_i when the block has large SNPs) of every block matrix; grid = (blocks, individual chunks)
__global__ void fill_test_rows_kernel(const BlockDesc* __restrict__ blocks, const uint8_t* __restrict__ tbed,
                                      int32_t tpitch, const int32_t* __restrict__ sel, int32_t n_test,
                                      const int32_t* __restrict__ tpos, const double* __restrict__ tmu,
                                      const double* __restrict__ tisd, double* __restrict__ sigma) {
    const BlockDesc bd = blocks[blockIdx.x];
    if (bd.m == 0) return;
    const int ml = bd.m - bd.ms;
    const int i0 = blockIdx.y * 16, i1 = min(n_test, i0 + 16);
    double* base = sigma + bd.moff + (size_t)(bd.mp + 8) * bd.ld;
    for (int i = i0; i < i1; ++i) {
        const int s = sel[i];
        double* rowA = base + (size_t)i * bd.ld;
        double* rowB = base + (size_t)(n_test + i) * bd.ld;
        for (int j = threadIdx.x; j < bd.ld; j += blockDim.x) {
            double x = 0.0;
            if (j < bd.m) {
                const int64_t g = bd.goff + j;
                const unsigned c = (tbed[(size_t)tpos[g] * tpitch + (s >> 2)] >> (2 * (s & 3))) & 3u;
                const double d = (c == 0u) ? 2.0 : (c == 2u) ? 1.0 : 0.0;
                x = (c == 1u) ? 0.0 : (d - tmu[g]) * tisd[g];
            }
            rowA[j] = (j < bd.ms) ? x : 0.0;
            if (ml > 0) rowB[j] = (j >= bd.ms) ? x : 0.0;
        }
    }
}

// one warp per (block, individual): the four row norms and the variance entry
__global__ void variance_kernel(const BlockDesc* __restrict__ blocks, int32_t n_blocks, int32_t n_test,
                                const double* __restrict__ sigma, const double* __restrict__ Lbuf,
                                double sigma_s, double n_obs, double* __restrict__ out) {
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= (int64_t)n_blocks * n_test) return;
    const int b = (int)(w / n_test), i = (int)(w - (int64_t)b * n_test);
    const int lane = threadIdx.x & 31;
    const BlockDesc bd = blocks[b];
    double d = 0.0;
    if (bd.m > 0) {
        const int ml = bd.m - bd.ms;
        const size_t rA = (size_t)bd.moff + (size_t)(bd.mp + 8 + i) * bd.ld;
        const size_t rB = (size_t)bd.moff + (size_t)(bd.mp + 8 + n_test + i) * bd.ld;
        double xs2 = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;
        for (int j = lane; j < bd.m; j += 32) {
            const double ya = Lbuf[rA + j];
            if (j < bd.ms) { const double x = sigma[rA + j]; xs2 += x * x; q1 += ya * ya; }
            else { q2 += ya * ya; if (ml > 0) { const double yb = Lbuf[rB + j]; q3 += yb * yb; } }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            xs2 += __shfl_xor_sync(0xffffffffu, xs2, o);
            q1 += __shfl_xor_sync(0xffffffffu, q1, o);
            q2 += __shfl_xor_sync(0xffffffffu, q2, o);
            q3 += __shfl_xor_sync(0xffffffffu, q3, o);
        }
        const double c = 1.0 / (sigma_s * n_obs);
        d = q3 / n_obs + sigma_s * (xs2 - c * q1 + c * q2);
    }
    if (lane == 0) out[w] = d;
}

cudaError_t launch_test_rows(const BlockDesc* blocks, int32_t n_blocks, const uint8_t* tbed, int32_t n_test_total,
                             const int32_t* sel, int32_t n_test, const int32_t* tpos, int64_t n_rows, double* tmu,
                             double* tisd, double* sigma, cudaStream_t st) {
    if (n_blocks == 0 || n_rows == 0) return cudaSuccess;
    const int32_t tpitch = (n_test_total + 3) / 4;
    test_stats_kernel<<<(unsigned)((n_rows + 7) / 8), 256, 0, st>>>(tbed, tpitch, sel, n_test, tpos, n_rows, tmu, tisd);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    dim3 grid(n_blocks, (n_test + 15) / 16);
    fill_test_rows_kernel<<<grid, 256, 0, st>>>(blocks, tbed, tpitch, sel, n_test, tpos, tmu, tisd, sigma);
    return cudaGetLastError();
}

cudaError_t launch_variance(const BlockDesc* blocks, int32_t n_blocks, int32_t n_test, const double* sigma,
                            const double* Lbuf, double sigma_s, double n_obs, double* out, cudaStream_t st) {
    const int64_t warps = (int64_t)n_blocks * n_test;
    if (warps == 0) return cudaSuccess;
    variance_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(blocks, n_blocks, n_test, sigma, Lbuf, sigma_s, n_obs, out);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Quadratic form of the reference's `valid` tool (scr/validate.cpp:255-258): deno_b = z' Sigma_b z from the LOWER
// triangle of Sigma_b (as the Gram kernels write it): sum_i z_i (Sigma_ii z_i + 2 sum_{j<i} Sigma_ij z_j).
// One CTA per block, one warp per row (coalesced row reads), fixed-order reductions.  HBM-bound: 4 m^2 bytes.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
quadform_kernel(const BlockDesc* __restrict__ blocks, const double* __restrict__ sigma, const double* __restrict__ z,
                double* __restrict__ out) {
    __shared__ double wsum[8];
    const BlockDesc bd = blocks[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const double* S = sigma + bd.moff;
    const double* zb = z + bd.goff;
    double acc = 0.0;                                  // lane 0 of each warp: sum over its rows
    for (int i = warp; i < bd.m; i += 8) {
        const double* row = S + (size_t)i * bd.ld;
        double p0 = 0.0, p1 = 0.0;
        int j = lane;
        for (; j + 32 < i; j += 64) { p0 += row[j] * zb[j]; p1 += row[j + 32] * zb[j + 32]; }
        if (j < i) p0 += row[j] * zb[j];
        double p = p0 + p1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
        const double zi = zb[i];
        acc += zi * (row[i] * zi + 2.0 * p);
    }
    if (lane == 0) wsum[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += wsum[w];
        out[blockIdx.x] = t;
    }
}

cudaError_t launch_quadform(const BlockDesc* blocks, int32_t n_blocks, const double* sigma, const double* z, double* out,
                            cudaStream_t st) {
    if (n_blocks == 0) return cudaSuccess;
    quadform_kernel<<<n_blocks, 256, 0, st>>>(blocks, sigma, z, out);
    return cudaGetLastError();
}

}  // namespace dbslmm
