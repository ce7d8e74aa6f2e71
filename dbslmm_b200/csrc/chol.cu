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
//   chol_diag_kernel  : potrf of the (already updated) 64x64 diagonal tile T_kk and W_kk = L_kk^-1;
//                       L_kk -> lower, W_kk^T -> upper triangle of the tile.
//   chol_panel_kernel : for each 128-row macro tile below: C = K_ik - L_i,0:k L_k,0:k^T (DMMA,
//                       cp.async 3-stage ring), L_ik = C W_kk^T (DMMA) -- TRSM as a GEMM -- and the
//                       look-ahead T_ii -= L_ik L_ik^T on the macro tile's own diagonal tiles.
// The z-scores ride along as matrix row `mp`, so that row of L ends up holding y = L^-1 z;
// backsolve_kernel then solves L^T x = y per block and writes beta = x / sqrt(N).
// Matrices are row-major, lower triangle, ld = mp (m padded to 8 with identity rows).
#include <algorithm>
#include <cooperative_groups.h>
#include "common.cuh"
#include "kernels.h"

namespace dbslmm {

static constexpr int NB = 64;          // panel width
static constexpr int TM = 128;         // rows per macro tile
static constexpr int KC = 16;          // K chunk per pipeline stage
static constexpr int LDS = KC + 8;     // padded smem row stride (doubles): conflict-free 128-bit fragment loads
static constexpr int NST = 3;          // cp.async stages
static constexpr int P_STAGE = TM * LDS;
static constexpr int Q_STAGE = NB * LDS;
static constexpr int STAGE = P_STAGE + Q_STAGE;
static constexpr int LDT = NB + 4;     // stride of 64-wide epilogue tiles
static constexpr int CHOL_THREADS = 256;
static constexpr int SMEM_PIPE = NST * STAGE * 8;
static constexpr int SMEM_CT = 8 * 16 * LDT * 8;                           // macro tile C / L, aliases the pipeline
static constexpr int W_OFF = SMEM_CT / 8;                                  // W tile follows the macro tile (doubles)
static constexpr int SMEM_EPI = SMEM_CT + NB * LDT * 8;
static constexpr int SMEM_CHOL = (SMEM_PIPE > SMEM_EPI ? SMEM_PIPE : SMEM_EPI);

// acc (16 rows x 64 cols per warp) += P[r0.., 0:K] * Q[q0.., 0:K]^T, both row-major with K contiguous.
// prow/qrow = number of valid rows (others are zero-filled).  All 256 threads must call.
__device__ __forceinline__ void gemm_nt_core(const double* __restrict__ Pg, const double* __restrict__ Qg, int ld,
                                             int prow, int qrow, int kbeg, int kend, double* smem,
                                             double (&acc)[2][8][2]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int nchunk = (kend - kbeg) / KC;     // K range [kbeg, kend), both multiples of KC
    const bool active = (16 * warp < prow);
    Pg += kbeg;
    Qg += kbeg;

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
            // One 128-bit load feeds two DMMAs: within each 8-wide K group lane t owns k = 2t (.x) and
            // k = 2t+1 (.y); A and B use the same assignment, so every product pairs the same k.
            const double* Ps = smem + (kc % NST) * STAGE + (16 * warp + g) * LDS + 2 * t;
            const double* Qs = smem + (kc % NST) * STAGE + P_STAGE + g * LDS + 2 * t;
#pragma unroll
            for (int s8 = 0; s8 < KC / 8; ++s8) {
                const double2 a0 = *reinterpret_cast<const double2*>(Ps + s8 * 8);
                const double2 a1 = *reinterpret_cast<const double2*>(Ps + 8 * LDS + s8 * 8);
                double2 b[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) b[c] = *reinterpret_cast<const double2*>(Qs + c * 8 * LDS + s8 * 8);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    if (8 * c < qrow) {
                        dmma884(acc[0][c][0], acc[0][c][1], a0.x, b[c].x);
                        dmma884(acc[1][c][0], acc[1][c][1], a1.x, b[c].x);
                        dmma884(acc[0][c][0], acc[0][c][1], a0.y, b[c].y);
                        dmma884(acc[1][c][0], acc[1][c][1], a1.y, b[c].y);
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Diagonal tile of panel k for every active block: factor + invert (no GEMM: the tile arrives
// fully updated, see chol_panel_kernel's look-ahead SYRK).
//   potrf: 4 sub-blocks of 16; each 16x16 diagonal sub-block is factored and inverted by ONE
//   warp in registers (shuffles), the sub-panel below and the trailing part by all 8 warps.
//   W = L^-1 is then assembled block-wise from the four 16x16 inverses.
// Tiles narrower than 64 (last panel) are padded with identity, so the code path is uniform.
// ------------------------------------------------------------------------------------------
static constexpr int DT = NB + 1;   // odd stride: conflict-free row and column walks in FP64
static constexpr int SMEM_DIAG = (2 * NB * DT + 3 * 16 * 17) * 8;

__device__ __forceinline__ void diag_body(const BlockDesc& bd, int blk, int k, const double* __restrict__ sigma,
                                          double* __restrict__ Lbuf, double* __restrict__ wbuf, double ridge,
                                          int32_t* __restrict__ status, double* smem) {
    double* T = smem;                        // [64][DT] tile, becomes L (lower)
    double* Wf = T + NB * DT;                // [64][DT] W = L^-1 (lower)
    double* tmp = Wf + NB * DT;              // [16][17]
    const int pc0 = k * NB;
    const int wk = min(NB, bd.mp - pc0);
    const int ld = bd.ld;
    double* Lb = Lbuf + bd.moff;
    const double* src = (k == 0 ? sigma : Lbuf) + bd.moff;     // step 0 reads Sigma, later steps the accumulated tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    {
        // all 16 loads of a thread are issued back to back (independent), then consumed
        double tv[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int idx = tid + u * CHOL_THREADS;
            const int a = idx >> 6, b = idx & 63;
            const bool ld_it = (a < wk) && (b <= a);
            tv[u] = ld_it ? __ldcg(src + (size_t)(pc0 + a) * ld + pc0 + b) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const int idx = tid + u * CHOL_THREADS;
            const int a = idx >> 6, b = idx & 63;
            double v = tv[u];
            if (a == b) {
                if (a >= wk) v = 1.0;                                   // identity padding
                else if (k == 0 && pc0 + a < bd.ms) v += ridge;
            }
            T[a * DT + b] = v;
            Wf[a * DT + b] = 0.0;
        }
    }
    __syncthreads();

    bool bad = false;
#pragma unroll 1
    for (int jb = 0; jb < 4; ++jb) {
        const int o = 16 * jb;
        if (warp == 0) {
            // ---- 16x16 potrf + inverse in registers; lane r (and r+16) holds row r
            const int r = lane & 15;
            double row[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) row[c] = T[(o + r) * DT + o + c];
            double dinv[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double d = __shfl_sync(0xffffffffu, row[j], j);
                if (!(d > 0.0)) bad = true;
                const double inv = rsqrt(d);
                const double sq = d * inv;
                dinv[j] = inv;
                double lrj = row[j] * inv;
                if (r == j) lrj = sq;
                if (r < j) lrj = 0.0;
                row[j] = lrj;
#pragma unroll
                for (int c = j + 1; c < 16; ++c) {
                    const double lcj = __shfl_sync(0xffffffffu, lrj, c);
                    row[c] -= lrj * lcj;
                }
            }
            if (lane < 16) {
#pragma unroll
                for (int c = 0; c < 16; ++c) T[(o + r) * DT + o + c] = (c <= r) ? row[c] : 0.0;
            }
            __syncwarp();
            // inverse: lane c owns column c of W16; w[i] = -(sum_{k<i} l_ik w[k]) / l_ii
            const int c = r;
            double w[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
                for (int kk = 0; kk + 1 < i; kk += 2) {
                    acc0 += T[(o + i) * DT + o + kk] * w[kk];
                    acc1 += T[(o + i) * DT + o + kk + 1] * w[kk + 1];
                }
                if (i & 1) acc0 += T[(o + i) * DT + o + i - 1] * w[i - 1];
                double wi = -(acc0 + acc1) * dinv[i];
                if (i == c) wi = dinv[i];
                if (i < c) wi = 0.0;
                w[i] = wi;
            }
            if (lane < 16) {
#pragma unroll
                for (int i = 0; i < 16; ++i) Wf[(o + i) * DT + o + c] = w[i];
            }
        }
        __syncthreads();
        if (jb == 3) break;
        // ---- sub-panel below: X = T[rows, o:o+16] * W16^T  (rows o+16 .. 63)
        const int nrem = 48 - o;
        double x[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int oi = tid + u * CHOL_THREADS;
            x[u] = 0.0;
            if (oi < nrem * 16) {
                const int i = o + 16 + (oi >> 4), c = oi & 15;
                double a = 0.0;
                for (int cp = 0; cp <= c; ++cp) a += T[i * DT + o + cp] * Wf[(o + c) * DT + o + cp];
                x[u] = a;
            }
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int oi = tid + u * CHOL_THREADS;
            if (oi < nrem * 16) T[(o + 16 + (oi >> 4)) * DT + o + (oi & 15)] = x[u];
        }
        __syncthreads();
        // ---- trailing update of the remaining lower triangle
        for (int oi = tid; oi < nrem * nrem; oi += CHOL_THREADS) {
            const int i = oi / nrem, c = oi - i * nrem;
            if (c <= i) {
                const double* xi = T + (o + 16 + i) * DT + o;
                const double* xc = T + (o + 16 + c) * DT + o;
                double a = 0.0;
#pragma unroll
                for (int kk = 0; kk < 16; ++kk) a += xi[kk] * xc[kk];
                T[(o + 16 + i) * DT + o + 16 + c] -= a;
            }
        }
        __syncthreads();
    }
    if (warp == 0 && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&status[blk], 1);
    // ---- off-diagonal 16x16 blocks of W:  W_ij = -W_ii * sum_{kb=j}^{i-1} L_i,kb W_kb,j, level by level in
    // the block distance d = i - j (all blocks of a level at once; tmp holds up to three 16x16 products)
    {
        const int r = tid >> 4, c = tid & 15;
#pragma unroll 1
        for (int d = 1; d < 4; ++d) {
            const int nblk = 4 - d;
            for (int q = 0; q < nblk; ++q) {
                const int i = d + q, j = q;
                double a = 0.0;
                for (int kk = 16 * j; kk < 16 * i; ++kk) a += T[(16 * i + r) * DT + kk] * Wf[kk * DT + 16 * j + c];
                tmp[q * 272 + r * 17 + c] = a;
            }
            __syncthreads();
            for (int q = 0; q < nblk; ++q) {
                const int i = d + q, j = q;
                double wv = 0.0;
                for (int kk = 0; kk <= r; ++kk) wv += Wf[(16 * i + r) * DT + 16 * i + kk] * tmp[q * 272 + kk * 17 + c];
                Wf[(16 * i + r) * DT + 16 * j + c] = -wv;
            }
            __syncthreads();
        }
    }
    // ---- write back: lower = L_kk, strict upper = W_kk^T (used by the back substitution) ...
    for (int idx = tid; idx < wk * wk; idx += CHOL_THREADS) {
        const int a = idx / wk, b = idx - a * wk;
        Lb[(size_t)(pc0 + a) * ld + pc0 + b] = (b <= a) ? T[a * DT + b] : Wf[b * DT + a];
    }
    // ... and W_kk as a dense 64x64 lower-triangular tile for the panel kernel of this step, which
    // streams it into shared memory with cp.async while its GEMM loop runs
    double* wb = wbuf + (size_t)blk * (NB * NB);
    for (int idx = tid; idx < NB * NB; idx += CHOL_THREADS) {
        const int a = idx >> 6, b = idx & 63;
        wb[idx] = (b <= a) ? Wf[a * DT + b] : 0.0;
    }
}

__global__ void __launch_bounds__(CHOL_THREADS, 3)
chol_diag_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ items, int32_t k,
                 const double* __restrict__ sigma, double* __restrict__ Lbuf, double* __restrict__ wbuf,
                 double ridge, int32_t* __restrict__ status) {
    extern __shared__ __align__(16) double smem[];
    const int blk = items[blockIdx.x];
    const BlockDesc bd = blocks[blk];
    diag_body(bd, blk, k, sigma, Lbuf, wbuf, ridge, status, smem);
}

// ------------------------------------------------------------------------------------------
// Rows below the diagonal tile of panel k:
//   C = K_ik - L_i,0:k L_k,0:k^T (DMMA, cp.async ring), L_ik = C W_kk^T (DMMA), then the
//   look-ahead: each of the (up to two) 64-row tiles of this macro tile immediately applies
//   its contribution  T_ii -= L_ik L_ik^T  to ITS OWN diagonal tile (one writer per tile and
//   step, so no atomics), so the diagonal kernel never runs a GEMM.
// ------------------------------------------------------------------------------------------
// item: x = block, y = macro tile, z = slice | nslices << 8, w = split group id.
// Returns false when this CTA was a non-final split-K slice (nothing more to do).
__device__ __forceinline__ bool panel_body(const BlockDesc& bd, const int4 item, int k, const double* __restrict__ sigma,
                                           double* __restrict__ Lbuf, const double* __restrict__ wbuf, double ridge,
                                           double* __restrict__ scratch, int32_t* __restrict__ counters,
                                           int32_t group_base, double* smem, int* s_last_p) {
    int& s_last = *s_last_p;
    const int pc0 = k * NB;
    const int wk = min(NB, bd.mp - pc0);
    const int ld = bd.ld;
    const int nrows = bd.nrows;
    const int r0 = pc0 + wk + item.y * TM;
    const int prow = min(TM, nrows - r0);
    double* Lb = Lbuf + bd.moff;
    const double* Sb = sigma + bd.moff;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

    const int slice = item.z & 0xFF, nsl = item.z >> 8;
    double* Wsm = smem + W_OFF;                   // [64][LDT], W[c][c'] = (L_kk^-1)[c][c'], zero above the diagonal
    // accumulators start at -K_ik (slice 0 only), so the Sigma tile's HBM latency hides behind the pipeline
    // prologue: after the loop acc = L_i,0:k L_k,0:k^T - K_ik = -C
    double acc[2][8][2];
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int rl = 16 * warp + 8 * f + g, cc = 8 * c + 2 * t;
            double2 a = make_double2(0.0, 0.0);
            if (slice == 0 && rl < prow && cc < wk) a = *reinterpret_cast<const double2*>(Sb + (size_t)(r0 + rl) * ld + pc0 + cc);
            acc[f][c][0] = -a.x;
            acc[f][c][1] = -a.y;
        }
    {
        // split-K: slice s of nsl owns 64-wide K blocks [k*s/nsl, k*(s+1)/nsl)
        const int kb = (k * slice) / nsl * NB, ke = (k * (slice + 1)) / nsl * NB;
        gemm_nt_core(Lb + (size_t)r0 * ld, Lb + (size_t)pc0 * ld, ld, prow, wk, kb, ke, smem, acc);
    }
    if (nsl > 1) {
        // partial sums go to scratch; the CTA that arrives last adds them up IN SLICE ORDER (deterministic)
        double* part = scratch + ((size_t)(item.w - group_base) * nsl) * (TM * NB);
        double2* mine = reinterpret_cast<double2*>(part + (size_t)slice * (TM * NB)) + tid;
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int c = 0; c < 8; ++c) mine[(f * 8 + c) * CHOL_THREADS] = make_double2(acc[f][c][0], acc[f][c][1]);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = (atomicAdd(&counters[item.w], 1) == nsl - 1);
        __syncthreads();
        if (!s_last) return false;
        __threadfence();
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[f][c][0] = acc[f][c][1] = 0.0;
        for (int sl = 0; sl < nsl; ++sl) {
            const double2* src2 = reinterpret_cast<const double2*>(part + (size_t)sl * (TM * NB)) + tid;
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const double2 v = __ldcg(src2 + (f * 8 + c) * CHOL_THREADS);
                    acc[f][c][0] += v.x;
                    acc[f][c][1] += v.y;
                }
        }
    }

    {
        // W tile of this step (written dense by the diagonal kernel): its L2 latency overlaps the C write-out
        const double* wsrc = wbuf + (size_t)item.x * (NB * NB);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = tid + u * CHOL_THREADS;
            const int row = idx >> 5, ch = (idx & 31) * 2;
            cp_async16(Wsm + row * LDT + ch, wsrc + row * NB + ch, true);
        }
        cp_async_commit();
    }
    double* Ct = smem;                            // [128][LDT] macro tile, 16 rows per warp
    double* Cw = Ct + warp * 16 * LDT;
    const double* W = Wsm;
    const bool active = (16 * warp < prow);
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int cc = 8 * c + 2 * t;
            Cw[(8 * f + g) * LDT + cc] = -acc[f][c][0];
            Cw[(8 * f + g) * LDT + cc + 1] = -acc[f][c][1];
            acc[f][c][0] = acc[f][c][1] = 0.0;
        }
    cp_async_wait<0>();
    __syncthreads();
    const double* Ca = Cw + g * LDT + t;
    const double* Wb = W + g * LDT + t;
    const int ns4 = wk / 4;
    if (active) {
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
    }
    __syncwarp();
    // L tile: to global and back into shared memory (operand of the look-ahead SYRK)
#pragma unroll
    for (int f = 0; f < 2; ++f)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int rl = 16 * warp + 8 * f + g, cc = 8 * c + 2 * t;
            if (rl < prow && cc < wk)
                *reinterpret_cast<double2*>(Lb + (size_t)(r0 + rl) * ld + pc0 + cc) =
                    make_double2(acc[f][c][0], acc[f][c][1]);
            Cw[(8 * f + g) * LDT + cc] = acc[f][c][0];
            Cw[(8 * f + g) * LDT + cc + 1] = acc[f][c][1];
            acc[f][c][0] = acc[f][c][1] = 0.0;
        }
    // ---- look-ahead: T_ii -= L_ik L_ik^T for the 64-row tile this warp belongs to
    const int grp = warp >> 2, wl = warp & 3;
    const int trow0 = r0 + 64 * grp;              // first global row of the 64-row tile
    const int tw = min(NB, bd.mp - trow0);        // tile extent (<= 0: z rows, no diagonal tile)
    const bool look = (trow0 < bd.mp && wk == NB && 16 * wl < tw);   // (a narrow last panel has no diagonal tile below)
    if (look) {
        // the old tile values go straight into the accumulators (negated): their latency overlaps the barrier
        const double* src = (k == 0 ? sigma : Lbuf) + bd.moff;    // first touch reads Sigma (+ ridge)
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int rl = 16 * wl + 8 * f + g, cc = 8 * c + 2 * t;   // within the tile
                if (c <= 2 * wl + 1 && rl < tw && cc <= rl) {
                    double2 v = __ldcg(reinterpret_cast<const double2*>(src + (size_t)(trow0 + rl) * ld + trow0 + cc));
                    if (k == 0) {
                        if (cc == rl && trow0 + rl < bd.ms) v.x += ridge;
                        if (cc + 1 == rl && trow0 + rl < bd.ms) v.y += ridge;   // odd rows: the diagonal is the pair's second element
                    }
                    acc[f][c][0] = -v.x;
                    acc[f][c][1] = -v.y;
                }
            }
    }
    __syncthreads();
    if (look) {
        const double* A = Ct + (64 * grp + 16 * wl + g) * LDT + t;
        const double* B = Ct + (64 * grp + g) * LDT + t;
#pragma unroll 4
        for (int s4 = 0; s4 < NB / 4; ++s4) {
            const double a0 = A[s4 * 4], a1 = A[8 * LDT + s4 * 4];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                if (c <= 2 * wl + 1) {                             // lower triangle of the tile only
                    const double b = B[c * 8 * LDT + s4 * 4];
                    dmma884(acc[0][c][0], acc[0][c][1], a0, b);
                    dmma884(acc[1][c][0], acc[1][c][1], a1, b);
                }
            }
        }
        // acc = L L^T - T_old  =>  T_new = -acc
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int rl = 16 * wl + 8 * f + g, cc = 8 * c + 2 * t;   // within the tile
                if (c <= 2 * wl + 1 && rl < tw && cc <= rl)
                    *reinterpret_cast<double2*>(Lb + (size_t)(trow0 + rl) * ld + trow0 + cc) =
                        make_double2(-acc[f][c][0], -acc[f][c][1]);
            }
    }
    return true;
}

__global__ void __launch_bounds__(CHOL_THREADS, 2)
chol_panel_kernel(const BlockDesc* __restrict__ blocks, const int4* __restrict__ items, int32_t k,
                  const double* __restrict__ sigma, double* __restrict__ Lbuf, const double* __restrict__ wbuf,
                  double ridge, double* __restrict__ scratch, int32_t* __restrict__ counters, int32_t group_base) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_last;
    const int4 item = items[blockIdx.x];
    const BlockDesc bd = blocks[item.x];
    panel_body(bd, item, k, sigma, Lbuf, wbuf, ridge, scratch, counters, group_base, smem, &s_last);
}

// ------------------------------------------------------------------------------------------
// Back substitution L^T x = y (y = matrix row mp), beta = x / sqrt(N).  One CTA per block.
// ------------------------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(NT)
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
    constexpr int NRG = NT / 32;    // row groups
    double* red = smem + mp;        // [NRG][64]
    double* v = red + NRG * NB;     // [64]
    const int tid = threadIdx.x;
    const int K = (mp + NB - 1) / NB;
    for (int k = K - 1; k >= 0; --k) {
        const int pc0 = k * NB, wk = min(NB, mp - pc0), below = pc0 + wk;
        {
            // 32 lanes x double2 cover the 64 panel columns of one row; 8 row groups, 4 rows in flight each
            const int rg = tid >> 5, c2 = (tid & 31) * 2;
            double2 p0 = make_double2(0.0, 0.0), p1 = p0, p2 = p0, p3 = p0;
            if (c2 < wk) {
                const double* col = Lb + pc0 + c2;
                int i = below + rg;
                for (; i + 3 * NRG < mp; i += 4 * NRG) {
                    const double2 a0 = *reinterpret_cast<const double2*>(col + (size_t)i * ld);
                    const double2 a1 = *reinterpret_cast<const double2*>(col + (size_t)(i + NRG) * ld);
                    const double2 a2 = *reinterpret_cast<const double2*>(col + (size_t)(i + 2 * NRG) * ld);
                    const double2 a3 = *reinterpret_cast<const double2*>(col + (size_t)(i + 3 * NRG) * ld);
                    const double x0 = x[i], x1 = x[i + NRG], x2 = x[i + 2 * NRG], x3 = x[i + 3 * NRG];
                    p0.x += a0.x * x0; p0.y += a0.y * x0;
                    p1.x += a1.x * x1; p1.y += a1.y * x1;
                    p2.x += a2.x * x2; p2.y += a2.y * x2;
                    p3.x += a3.x * x3; p3.y += a3.y * x3;
                }
                for (; i < mp; i += NRG) {
                    const double2 a0 = *reinterpret_cast<const double2*>(col + (size_t)i * ld);
                    p0.x += a0.x * x[i]; p0.y += a0.y * x[i];
                }
            }
            red[rg * NB + c2] = (p0.x + p1.x) + (p2.x + p3.x);
            red[rg * NB + c2 + 1] = (p0.y + p1.y) + (p2.y + p3.y);
        }
        __syncthreads();
        if (tid < NB) {
            double sacc = 0.0;
#pragma unroll
            for (int q = 0; q < NRG; ++q) sacc += red[q * NB + tid];
            v[tid] = (tid < wk) ? y[pc0 + tid] - sacc : 0.0;
        }
        __syncthreads();
        if (tid < 256) {
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
    for (int j = tid; j < bd.m; j += NT) {
        const double b = x[j] * inv_sqrt_n;
        if (j < bd.ms) beta_s[bd.out_s + j] = b;
        else beta_l[bd.out_l + (j - bd.ms)] = b;
    }
}

// ------------------------------------------------------------------------------------------
// Back substitution for big blocks (mp > 1024): a thread-block CLUSTER of 8 CTAs per block.
// Each CTA streams one eighth of the rows below the current panel, the eight 64-entry partial
// sums are exchanged through distributed shared memory, and every CTA then forms x_k itself
// (redundantly: cheaper than a broadcast).  Two cluster barriers per panel.  One CTA alone is
// latency-bound at ~30 GB/s on a 36 MB factor and was the critical path of the 8-GPU shards.
// ------------------------------------------------------------------------------------------
static constexpr int kBsCluster = 8;
static constexpr int kBsThreads = 512;

__global__ void __cluster_dims__(kBsCluster, 1, 1) __launch_bounds__(kBsThreads)
backsolve_cluster_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ order,
                         const double* __restrict__ Lbuf, double inv_sqrt_n, double* __restrict__ beta_s,
                         double* __restrict__ beta_l) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double smem[];
    const unsigned cr = cluster.block_rank();
    const BlockDesc bd = blocks[order[blockIdx.x / kBsCluster]];
    const int mp = bd.mp, ld = bd.ld;
    const double* Lb = Lbuf + bd.moff;
    const double* y = Lb + (size_t)mp * ld;
    constexpr int NRG = kBsThreads / 32;          // row groups per CTA
    constexpr int RS = NRG * kBsCluster;          // row stride of one (CTA, row group)
    double* x = smem;                              // [mp]   full solution vector, kept by every CTA
    double* part = smem + mp;                      // [64]   this CTA's partial sums (read by the others)
    double* red = part + NB;                       // [NRG][64]
    double* v = red + NRG * NB;                    // [64]
    const int tid = threadIdx.x;
    const int K = (mp + NB - 1) / NB;
    for (int k = K - 1; k >= 0; --k) {
        const int pc0 = k * NB, wk = min(NB, mp - pc0), below = pc0 + wk;
        {
            const int rg = tid >> 5, c2 = (tid & 31) * 2;
            double2 p0 = make_double2(0.0, 0.0), p1 = p0, p2 = p0, p3 = p0;
            if (c2 < wk) {
                const double* col = Lb + pc0 + c2;
                int i = below + (int)cr * NRG + rg;
                for (; i + 3 * RS < mp; i += 4 * RS) {
                    const double2 a0 = *reinterpret_cast<const double2*>(col + (size_t)i * ld);
                    const double2 a1 = *reinterpret_cast<const double2*>(col + (size_t)(i + RS) * ld);
                    const double2 a2 = *reinterpret_cast<const double2*>(col + (size_t)(i + 2 * RS) * ld);
                    const double2 a3 = *reinterpret_cast<const double2*>(col + (size_t)(i + 3 * RS) * ld);
                    const double x0 = x[i], x1 = x[i + RS], x2 = x[i + 2 * RS], x3 = x[i + 3 * RS];
                    p0.x += a0.x * x0; p0.y += a0.y * x0;
                    p1.x += a1.x * x1; p1.y += a1.y * x1;
                    p2.x += a2.x * x2; p2.y += a2.y * x2;
                    p3.x += a3.x * x3; p3.y += a3.y * x3;
                }
                for (; i < mp; i += RS) {
                    const double2 a0 = *reinterpret_cast<const double2*>(col + (size_t)i * ld);
                    p0.x += a0.x * x[i]; p0.y += a0.y * x[i];
                }
            }
            red[rg * NB + c2] = (p0.x + p1.x) + (p2.x + p3.x);
            red[rg * NB + c2 + 1] = (p0.y + p1.y) + (p2.y + p3.y);
        }
        __syncthreads();
        if (tid < NB) {
            double sacc = 0.0;
#pragma unroll
            for (int q = 0; q < NRG; ++q) sacc += red[q * NB + tid];
            part[tid] = sacc;
        }
        cluster.sync();                            // all eight partial vectors are in place
        if (tid < NB) {
            double sacc = 0.0;
#pragma unroll
            for (unsigned r = 0; r < (unsigned)kBsCluster; ++r) sacc += cluster.map_shared_rank(part, r)[tid];   // fixed order
            v[tid] = (tid < wk) ? y[pc0 + tid] - sacc : 0.0;
        }
        __syncthreads();
        if (tid < 256) {
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
        cluster.sync();                            // everyone has read the partials: they may be overwritten
    }
    if (cr == 0) {
        for (int j = tid; j < bd.m; j += kBsThreads) {
            const double b = x[j] * inv_sqrt_n;
            if (j < bd.ms) beta_s[bd.out_s + j] = b;
            else beta_l[bd.out_l + (j - bd.ms)] = b;
        }
    }
}

cudaError_t chol_configure() {
    cudaError_t e = cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DIAG);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CHOL);
}

cudaError_t launch_chol_diag(const BlockDesc* blocks, const int32_t* items, int32_t n_items, int32_t k,
                             const double* sigma, double* L, double* wbuf, double ridge, int32_t* status,
                             cudaStream_t st) {
    if (n_items == 0) return cudaSuccess;
    chol_diag_kernel<<<n_items, CHOL_THREADS, SMEM_DIAG, st>>>(blocks, items, k, sigma, L, wbuf, ridge, status);
    return cudaGetLastError();
}
cudaError_t launch_chol_panel(const BlockDesc* blocks, const int4* items, int32_t n_items, int32_t k,
                              const double* sigma, double* L, const double* wbuf, double ridge, double* scratch,
                              int32_t* counters, int32_t group_base, cudaStream_t st) {
    if (n_items == 0) return cudaSuccess;
    chol_panel_kernel<<<n_items, CHOL_THREADS, SMEM_CHOL, st>>>(blocks, items, k, sigma, L, wbuf, ridge, scratch,
                                                                 counters, group_base);
    return cudaGetLastError();
}
// Back substitution of `n_blocks` blocks listed in `order`: an 8-CTA cluster per block for the big size
// classes (mp > 1024), one 256-thread CTA per block otherwise.
cudaError_t launch_backsolve(const BlockDesc* blocks, const int32_t* order, int32_t n_blocks, bool big,
                             const double* L, double inv_sqrt_n, double* beta_s, double* beta_l, int32_t max_mp,
                             cudaStream_t st) {
    if (n_blocks == 0) return cudaSuccess;
    if (big) {
        const size_t smem = (size_t)(max_mp + (kBsThreads / 32 + 2) * NB) * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(backsolve_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        backsolve_cluster_kernel<<<n_blocks * kBsCluster, kBsThreads, smem, st>>>(blocks, order, L, inv_sqrt_n, beta_s, beta_l);
    } else {
        const size_t smem = (size_t)(1024 + 9 * NB) * sizeof(double);
        backsolve_kernel<256><<<n_blocks, 256, smem, st>>>(blocks, order, L, inv_sqrt_n, beta_s, beta_l);
    }
    return cudaGetLastError();
}

}  // namespace dbslmm
