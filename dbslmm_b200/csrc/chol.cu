// chol.cu -- block solver (K4/K5): batched, variable-size FP64 Cholesky of the bordered
// block system, FP64 tensor-core (DMMA) panel updates, fused forward substitution.
//
// What it replaces: the (m_l + 2) Jacobi-PCG solves, the Schur complement and the GEMV
// glue of DBSLMMFIT::estBlock (reference scr/dbslmmfit.cpp:712-729, 759-764) and
// PCGv/PCGm (:629-678).  With the block's SNPs ordered small-first, the reference's
//     A = Sigma_ss + c I,  W = A^-1 Sigma_sl,  S = Sigma_ll - Sigma_ls W,
//     beta_l = S^-1 (z_l - Sigma_ls A^-1 z_s) / sqrt(N),  beta_s = (A^-1 z_s - sqrt(N) W beta_l) / sqrt(N)
// is exactly block elimination of the ONE symmetric positive definite system
//     K x = z,   K = Sigma + c * diag(1_small, 0_large),   beta = x / sqrt(N)
// (the Schur complement S is the trailing block of K's Cholesky factor), so every block --
// LMM or DBSLMM mode -- is one factorisation and one right-hand side.
//
// Algorithm: left-looking tile Cholesky, panel width 64, batched over ALL blocks of a size
// class per panel step k (two launches per step):
//   chol_diag_kernel  : T_kk = K_kk - L_k,0:k L_k,0:k^T (DMMA), potrf(T_kk) in shared memory,
//                       W_kk = L_kk^-1; L_kk -> lower, W_kk^T -> upper triangle of the tile.
//   chol_panel_kernel : for each 128-row macro tile below: C = K_ik - L_i,0:k L_k,0:k^T (DMMA,
//                       cp.async 3-stage ring), then L_ik = C W_kk^T (DMMA) -- TRSM as a GEMM.
// The z-scores ride along as matrix row `mp`, so that row of L ends up holding y = L^-1 z;
// backsolve_kernel then solves L^T x = y per block and writes beta = x / sqrt(N).
// Matrices are row-major, lower triangle, ld = mp (m padded to 8 with identity rows).
#include "common.cuh"
#include "kernels.h"

namespace dbslmm {

static constexpr int NB = 64;          // panel width
static constexpr int TM = 128;         // rows per macro tile
static constexpr int KC = 16;          // K chunk per pipeline stage
static constexpr int LDS = KC + 4;     // padded smem row stride (doubles): conflict-free 8x4 fragment loads
static constexpr int NST = 3;          // cp.async stages
static constexpr int P_STAGE = TM * LDS;
static constexpr int Q_STAGE = NB * LDS;
static constexpr int STAGE = P_STAGE + Q_STAGE;
static constexpr int LDT = NB + 4;     // stride of 64-wide epilogue tiles
static constexpr int CHOL_THREADS = 256;
static constexpr int SMEM_PIPE = NST * STAGE * 8;
static constexpr int SMEM_EPI = (8 * 16 * LDT + NB * LDT) * 8;
static constexpr int SMEM_CHOL = (SMEM_PIPE > SMEM_EPI ? SMEM_PIPE : SMEM_EPI);

// acc (16 rows x 64 cols per warp) += P[r0.., 0:K] * Q[q0.., 0:K]^T, both row-major with K contiguous.
// prow/qrow = number of valid rows (others are zero-filled).  All 256 threads must call.
__device__ __forceinline__ void gemm_nt_core(const double* __restrict__ Pg, const double* __restrict__ Qg, int ld,
                                             int prow, int qrow, int Kdim, double* smem,
                                             double (&acc)[2][8][2]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int nchunk = Kdim / KC;
    const bool active = (16 * warp < prow);

    auto load_stage = [&](int kc, int s) {
        double* Ps = smem + s * STAGE;
        double* Qs = Ps + P_STAGE;
        const int k0 = kc * KC;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int id = tid + u * CHOL_THREADS;
            const int r = id >> 3, c2 = (id & 7) * 2;
            const bool v = r < prow;
            cp_async16(Ps + r * LDS + c2, Pg + (size_t)(v ? r : 0) * ld + k0 + c2, v);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int id = tid + u * CHOL_THREADS;
            const int r = id >> 3, c2 = (id & 7) * 2;
            const bool v = r < qrow;
            cp_async16(Qs + r * LDS + c2, Qg + (size_t)(v ? r : 0) * ld + k0 + c2, v);
        }
    };

#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
        if (s < nchunk) load_stage(s, s);
        cp_async_commit();
    }
    for (int kc = 0; kc < nchunk; ++kc) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        const int nx = kc + NST - 1;
        if (nx < nchunk) load_stage(nx, nx % NST);
        cp_async_commit();
        if (active) {
            const double* Ps = smem + (kc % NST) * STAGE + (16 * warp + g) * LDS + t;
            const double* Qs = smem + (kc % NST) * STAGE + P_STAGE + g * LDS + t;
#pragma unroll
            for (int s4 = 0; s4 < KC / 4; ++s4) {
                const double a0 = Ps[s4 * 4], a1 = Ps[8 * LDS + s4 * 4];
                double b[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) b[c] = Qs[c * 8 * LDS + s4 * 4];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (8 * c < qrow) {
                        dmma884(acc[0][c][0], acc[0][c][1], a0, b[c]);
                        dmma884(acc[1][c][0], acc[1][c][1], a1, b[c]);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Diagonal tile of panel k for every active block: update, factor, invert.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CHOL_THREADS, 2)
chol_diag_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ items, int32_t k,
                 const double* __restrict__ sigma, double* __restrict__ Lbuf, double ridge,
                 int32_t* __restrict__ status) {
    extern __shared__ __align__(16) double smem[];
    const int blk = items[blockIdx.x];
    const BlockDesc bd = blocks[blk];
    const int pc0 = k * NB;
    const int wk = min(NB, bd.mp - pc0);
    const int ld = bd.ld;
    double* Lb = Lbuf + bd.moff;
    const double* Sb = sigma + bd.moff;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

    double acc[2][8][2];
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[f][c][0] = acc[f][c][1] = 0.0;
    const double* Qg = Lb + (size_t)pc0 * ld;
    gemm_nt_core(Qg, Qg, ld, wk, wk, pc0, smem, acc);

    double* T = smem;                 // [64][LDT]
    double* W = smem + NB * LDT;      // [64][LDT]
    for (int i = tid; i < NB * LDT; i += CHOL_THREADS) W[i] = 0.0;
    if (16 * warp < wk) {
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int r = 16 * warp + 8 * f + g, cc = 8 * c + 2 * t;
                if (r < wk && cc <= r) {                       // lower triangle (+ the pair partner)
                    const double2 a = *reinterpret_cast<const double2*>(Sb + (size_t)(pc0 + r) * ld + pc0 + cc);
                    double v0 = a.x - acc[f][c][0], v1 = a.y - acc[f][c][1];
                    const int gr = pc0 + r;
                    if (cc == r && gr < bd.ms) v0 += ridge;
                    if (cc + 1 == r && gr < bd.ms) v1 += ridge;
                    T[r * LDT + cc] = v0;
                    T[r * LDT + cc + 1] = v1;                  // (r, cc+1) may be above the diagonal: unused
                }
            }
    }
    // ---- potrf (right-looking, in shared memory)
    bool bad = false;
    const int ty = tid >> 4, tx = tid & 15;
    for (int j = 0; j < wk; ++j) {
        __syncthreads();
        const double d = T[j * LDT + j];
        if (!(d > 0.0)) bad = true;
        const double s = sqrt(d), inv = 1.0 / s;
        if (tid > j && tid < wk) T[tid * LDT + j] *= inv;
        __syncthreads();
        if (tid == 0) T[j * LDT + j] = s;
        for (int i = j + 1 + ty; i < wk; i += 16) {
            const double lij = T[i * LDT + j];
            for (int c = j + 1 + tx; c <= i; c += 16) T[i * LDT + c] -= lij * T[c * LDT + j];
        }
    }
    __syncthreads();
    if (bad && tid == 0) atomicOr(&status[blk], 1);
    // ---- W = L^-1 (lower), four lanes per column
    {
        const int c = tid >> 2, q = tid & 3;
        const int cw0 = (warp * 8);                            // first column handled by this warp
        if (q == 0 && c < wk) W[c * LDT + c] = 1.0 / T[c * LDT + c];
        __syncwarp();
        for (int i = cw0 + 1; i < wk; ++i) {
            double p = 0.0;
            if (c < wk && i > c)
                for (int kk = c + q; kk < i; kk += 4) p += T[i * LDT + kk] * W[kk * LDT + c];
            p += __shfl_xor_sync(0xffffffffu, p, 1);
            p += __shfl_xor_sync(0xffffffffu, p, 2);
            if (q == 0 && c < wk && i > c) W[i * LDT + c] = -p / T[i * LDT + i];
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- write back: lower = L_kk, strict upper = W_kk^T
    for (int idx = tid; idx < wk * wk; idx += CHOL_THREADS) {
        const int a = idx / wk, b = idx - a * wk;
        Lb[(size_t)(pc0 + a) * ld + pc0 + b] = (b <= a) ? T[a * LDT + b] : W[b * LDT + a];
    }
}

// ------------------------------------------------------------------------------------------
// Rows below the diagonal tile of panel k: update + triangular solve (as GEMM with W_kk^T).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CHOL_THREADS, 2)
chol_panel_kernel(const BlockDesc* __restrict__ blocks, const int2* __restrict__ items, int32_t k,
                  const double* __restrict__ sigma, double* __restrict__ Lbuf) {
    extern __shared__ __align__(16) double smem[];
    const int2 item = items[blockIdx.x];
    const BlockDesc bd = blocks[item.x];
    const int pc0 = k * NB;
    const int wk = min(NB, bd.mp - pc0);
    const int ld = bd.ld;
    const int nrows = bd.mp + 8;
    const int r0 = pc0 + wk + item.y * TM;
    const int prow = min(TM, nrows - r0);
    double* Lb = Lbuf + bd.moff;
    const double* Sb = sigma + bd.moff;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

    double acc[2][8][2];
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[f][c][0] = acc[f][c][1] = 0.0;
    gemm_nt_core(Lb + (size_t)r0 * ld, Lb + (size_t)pc0 * ld, ld, prow, wk, pc0, smem, acc);

    double* Cw = smem + warp * 16 * LDT;          // warp-private [16][LDT]
    double* W = smem + 8 * 16 * LDT;              // [64][LDT], W[c][c'] = (L_kk^-1)[c][c'], zero above diagonal
    for (int i = tid; i < NB * LDT; i += CHOL_THREADS) W[i] = 0.0;
    __syncthreads();
    for (int idx = tid; idx < wk * wk; idx += CHOL_THREADS) {
        const int a = idx / wk, b = idx - a * wk;
        const double v = Lb[(size_t)(pc0 + a) * ld + pc0 + b];
        if (b > a) W[b * LDT + a] = v;            // upper triangle stores W^T
        else if (b == a) W[a * LDT + a] = 1.0 / v;
    }
    const bool active = (16 * warp < prow);
    if (active) {
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int rl = 16 * warp + 8 * f + g, cc = 8 * c + 2 * t;
                double2 a = make_double2(0.0, 0.0);
                if (rl < prow && cc < wk) a = *reinterpret_cast<const double2*>(Sb + (size_t)(r0 + rl) * ld + pc0 + cc);
                Cw[(8 * f + g) * LDT + cc] = a.x - acc[f][c][0];
                Cw[(8 * f + g) * LDT + cc + 1] = a.y - acc[f][c][1];
                acc[f][c][0] = acc[f][c][1] = 0.0;
            }
    }
    __syncthreads();
    if (!active) return;
    const double* Ca = Cw + g * LDT + t;
    const double* Wb = W + g * LDT + t;
    const int ns4 = wk / 4;
#pragma unroll 4
    for (int s4 = 0; s4 < ns4; ++s4) {
        const double a0 = Ca[s4 * 4], a1 = Ca[8 * LDT + s4 * 4];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (4 * s4 <= 8 * c + 7 && 8 * c < wk) {          // W is lower triangular
                const double b = Wb[c * 8 * LDT + s4 * 4];
                dmma884(acc[0][c][0], acc[0][c][1], a0, b);
                dmma884(acc[1][c][0], acc[1][c][1], a1, b);
            }
        }
    }
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int rl = 16 * warp + 8 * f + g, cc = 8 * c + 2 * t;
            if (rl < prow && cc < wk)
                *reinterpret_cast<double2*>(Lb + (size_t)(r0 + rl) * ld + pc0 + cc) =
                    make_double2(acc[f][c][0], acc[f][c][1]);
        }
}

// ------------------------------------------------------------------------------------------
// Back substitution L^T x = y (y = matrix row mp), beta = x / sqrt(N).  One CTA per block.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(CHOL_THREADS)
backsolve_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ order,
                 const double* __restrict__ Lbuf, double inv_sqrt_n, double* __restrict__ beta_s,
                 double* __restrict__ beta_l) {
    extern __shared__ __align__(16) double smem[];
    const BlockDesc bd = blocks[order[blockIdx.x]];
    if (bd.m == 0) return;
    const int mp = bd.mp, ld = bd.ld;
    const double* Lb = Lbuf + bd.moff;
    const double* y = Lb + (size_t)mp * ld;
    double* x = smem;               // [mp]
    double* red = smem + mp;        // [4][64]
    double* v = red + 4 * NB;       // [64]
    const int tid = threadIdx.x;
    const int K = (mp + NB - 1) / NB;
    for (int k = K - 1; k >= 0; --k) {
        const int pc0 = k * NB, wk = min(NB, mp - pc0), below = pc0 + wk;
        {
            const int rg = tid >> 6, c = tid & 63;
            double p0 = 0.0, p1 = 0.0;
            if (c < wk) {
                int i = below + rg;
                for (; i + 4 < mp; i += 8) {
                    p0 += Lb[(size_t)i * ld + pc0 + c] * x[i];
                    p1 += Lb[(size_t)(i + 4) * ld + pc0 + c] * x[i + 4];
                }
                if (i < mp) p0 += Lb[(size_t)i * ld + pc0 + c] * x[i];
            }
            red[rg * NB + c] = p0 + p1;
        }
        __syncthreads();
        if (tid < NB) v[tid] = (tid < wk) ? y[pc0 + tid] - (red[tid] + red[NB + tid] + red[2 * NB + tid] + red[3 * NB + tid]) : 0.0;
        __syncthreads();
        {
            const int c = tid >> 2, q = tid & 3;
            double p = 0.0;
            if (c < wk) {
                const double* row = Lb + (size_t)(pc0 + c) * ld + pc0;
                for (int cp = c + q; cp < wk; cp += 4) p += ((cp == c) ? 1.0 / row[c] : row[cp]) * v[cp];
            }
            p += __shfl_xor_sync(0xffffffffu, p, 1);
            p += __shfl_xor_sync(0xffffffffu, p, 2);
            if (q == 0 && c < wk) x[pc0 + c] = p;
        }
        __syncthreads();
    }
    for (int j = tid; j < bd.m; j += CHOL_THREADS) {
        const double b = x[j] * inv_sqrt_n;
        if (j < bd.ms) beta_s[bd.out_s + j] = b;
        else beta_l[bd.out_l + (j - bd.ms)] = b;
    }
}

cudaError_t chol_configure() {
    cudaError_t e = cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CHOL);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CHOL);
}

cudaError_t launch_chol_diag(const BlockDesc* blocks, const int32_t* items, int32_t n_items, int32_t k,
                             const double* sigma, double* L, double ridge, int32_t* status, cudaStream_t st) {
    if (n_items == 0) return cudaSuccess;
    chol_diag_kernel<<<n_items, CHOL_THREADS, SMEM_CHOL, st>>>(blocks, items, k, sigma, L, ridge, status);
    return cudaGetLastError();
}
cudaError_t launch_chol_panel(const BlockDesc* blocks, const int2* items, int32_t n_items, int32_t k,
                              const double* sigma, double* L, cudaStream_t st) {
    if (n_items == 0) return cudaSuccess;
    chol_panel_kernel<<<n_items, CHOL_THREADS, SMEM_CHOL, st>>>(blocks, items, k, sigma, L);
    return cudaGetLastError();
}
cudaError_t launch_backsolve(const BlockDesc* blocks, const int32_t* order, int32_t n_blocks, const double* L,
                             double inv_sqrt_n, double* beta_s, double* beta_l, int32_t max_mp, cudaStream_t st) {
    if (n_blocks == 0) return cudaSuccess;
    const size_t smem = (size_t)(max_mp + 5 * NB) * sizeof(double);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(backsolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    backsolve_kernel<<<n_blocks, CHOL_THREADS, smem, st>>>(blocks, order, L, inv_sqrt_n, beta_s, beta_l);
    return cudaGetLastError();
}

}  // namespace dbslmm
