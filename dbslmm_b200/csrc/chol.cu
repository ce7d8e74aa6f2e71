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
// Algorithm: left-looking tile Cholesky, panel width 64, batched over ALL blocks of a batch (a size class, or an
// upload region of a streaming fit) per panel step k -- ONE launch per step:
//   chol_panel_kernel : for each 128-row macro tile below the diagonal tile: -C = L_i,0:k L_k,0:k^T - K_ik (DMMA,
//                       cp.async 3-stage ring), -L_ik = (-C) W_kk^T (DMMA, TRSM as a GEMM fed from the accumulator
//                       fragments), the look-ahead T_ii -= L_ik L_ik^T on the macro tile's own diagonal tiles, and the
//                       factorisation of the diagonal tile of panel k+1 (diag_body): by the CTA of macro tile 0 at the
//                       end of a multi-wave step, or by extra CTAs at the head of step k+1 when that step is
//                       chain-bound (they overlap its main loops and publish W through a flag).
//   chol_diag_kernel  : diag_body on its own, for k = 0: potrf of the 64x64 diagonal tile T_kk and W_kk = L_kk^-1;
//                       L_kk -> lower, W_kk^T -> upper triangle of the tile, W_kk dense for the panel kernel.
// The z-scores ride along as matrix row `mp`, so that row of L ends up holding y = L^-1 z;
// backsolve_kernel then solves L^T x = y per block and writes beta = x / sqrt(N).
// Matrices are row-major, lower triangle, ld = mp (m padded to 8 with identity rows).
#include <algorithm>
#include <cstdlib>
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
static constexpr int LDW = NB + 8;     // stride of the 64-wide epilogue tiles (W, L halves): conflict-free 128-bit fragment accesses
static constexpr int CHOL_THREADS = 256;
static constexpr int SMEM_PIPE = NST * STAGE * 8;
// The epilogue re-uses the three pipeline stages: one holds W_kk, the other two the two 64-row halves of L_ik.
static_assert(NB * LDW == STAGE, "a 64 x LDW epilogue tile must fill exactly one pipeline stage");
static constexpr int SMEM_CHOL = SMEM_PIPE;

// W_kk = L_kk^-1 travels from the diagonal-tile code to the panel CTAs as a ready-made shared-memory IMAGE: the ten
// 16x16 blocks of its lower triangle (block (rb, cb), cb <= rb, at 2 KB * (rb (rb + 1) / 2 + cb)), each block in the
// format a 128-byte-swizzled, row-permuting TMA box would have produced (8-row groups of 1 KB, rows of a group in the
// order 0,2,4,6,1,3,5,7, 16-byte chunks XOR-ed with the row slot).  One 1-D bulk copy of 20 KB brings it in, and the
// TRSM reads it with the same conflict-free 128-bit fragment loads as the ring boxes.
static constexpr int W_IMG_DOUBLES = 10 * 256;
static constexpr int W_IMG_BYTES = W_IMG_DOUBLES * 8;
__host__ __device__ __forceinline__ int w_img_off(int a, int b) {       // element (row a, column b), (b >> 4) <= (a >> 4)
    const int rb = a >> 4, cb = b >> 4;
    const int g = a & 7, sig = (g >> 1) | ((g & 1) << 2);
    return (rb * (rb + 1) / 2 + cb) * 256 + ((a >> 3) & 1) * 128 + sig * 16 + ((((b & 15) >> 1) ^ sig) << 1) + (b & 1);
}

// acc (16 rows x 64 cols per warp) += P[r0.., 0:K] * Q[q0.., 0:K]^T, both row-major with K contiguous.
// prow/qrow = number of valid rows (others are zero-filled).  `wrow` = this warp's 16-row slot of the macro tile.
// The dense 64x64 tile `wsrc` (W_kk of this step) rides through the ring as the chunk AFTER the last K chunk, so it
// has landed (stride LDW) in stage (nchunk % NST) when the loop ends, at no extra latency.  All 256 threads must call.
// `init_acc` runs right after the ring prologue has been issued: the caller's accumulator preload (global loads)
// then overlaps the first stages' latency instead of preceding it.
template <class InitAcc>
__device__ __forceinline__ void gemm_nt_core(const double* __restrict__ Pg, const double* __restrict__ Qg, int ld,
                                             int prow, int qrow, int kbeg, int kend, const double* __restrict__ wsrc,
                                             int wrow, double* smem, double (&acc)[2][8][2], InitAcc&& init_acc) {
    const int tid = threadIdx.x, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int nchunk = (kend - kbeg) / KC;     // K range [kbeg, kend), both multiples of KC
    const bool active = (16 * wrow < prow);
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
    auto load_w = [&](int s) {
        double* Ws = smem + s * STAGE;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = tid + u * CHOL_THREADS;
            const int row = idx >> 5, ch = (idx & 31) * 2;
            const bool v = (ch >> 4) <= (row >> 4);                  // 16x16 blocks above the diagonal: zero-filled
            cp_async16(Ws + row * LDW + ch, wsrc + (v ? w_img_off(row, ch) : 0), v);
        }
    };

#pragma unroll
    for (int s = 0; s < NST - 1; ++s) {
        if (s < nchunk) load_stage(s, s);
        else if (s == nchunk && wsrc != nullptr) load_w(s);
        cp_async_commit();
    }
    init_acc();
    for (int kc = 0; kc < nchunk; ++kc) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        const int nx = kc + NST - 1;
        if (nx < nchunk) load_stage(nx, nx % NST);
        else if (nx == nchunk && wsrc != nullptr) load_w(nx % NST);
        cp_async_commit();
        if (active) {
            // One 128-bit load feeds two DMMAs: within each 8-wide K group lane t owns k = 2t (.x) and
            // k = 2t+1 (.y); A and B use the same assignment, so every product pairs the same k.
            const double* Ps = smem + (kc % NST) * STAGE + (16 * wrow + g) * LDS + 2 * t;
            const double* Qs = smem + (kc % NST) * STAGE + P_STAGE + g * LDS + 2 * t;
#pragma unroll
            for (int s8 = 0; s8 < KC / 8; ++s8) {
                const double2 a0 = *reinterpret_cast<const double2*>(Ps + s8 * 8);
                const double2 a1 = *reinterpret_cast<const double2*>(Ps + 8 * LDS + s8 * 8);
                double2 b[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) b[c] = *reinterpret_cast<const double2*>(Qs + c * 8 * LDS + s8 * 8);
                // no per-column predicate (Q rows >= qrow are zero-filled): a predicated mma.sync costs a
                // WARPSYNC + NOP pair per DMMA in SASS
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    dmma884(acc[0][c][0], acc[0][c][1], a0.x, b[c].x);
                    dmma884(acc[1][c][0], acc[1][c][1], a1.x, b[c].x);
                    dmma884(acc[0][c][0], acc[0][c][1], a0.y, b[c].y);
                    dmma884(acc[1][c][0], acc[1][c][1], a1.y, b[c].y);
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

// ------------------------------------------------------------------------------------------
// Diagonal tile of panel k for every active block: factor + invert (no GEMM: the tile arrives
// fully updated, see chol_panel_kernel's look-ahead SYRK).  This is a latency chain (64 dependent
// column eliminations), so everything that is not on the chain is moved off it:
//   * warp 0 factors one 16x16 diagonal sub-block at a time in registers (shuffles);
//   * the 16-column sub-panel below is solved by forward substitution, one row per thread
//     (warps 2-4), while warp 1 inverts the 16x16 factor;
//   * warp 0 then updates only the NEXT 16x16 diagonal sub-block and goes straight on with its
//     factorisation, while warps 1-7 update the rest of the trailing matrix in its shadow;
//   * W = L^-1 is assembled from the four 16x16 inverses by two levels of 2x2 block inversion.
// Tiles narrower than 64 (last panel) are padded with identity, so the code path is uniform.
// ------------------------------------------------------------------------------------------
#ifndef DIAG_STAMP
#define DIAG_STAMP(n)               // probe hook (tools/diag_probe.cu records clock64() here)
#endif
static constexpr int DT = NB + 1;   // odd stride: conflict-free row and column walks in FP64
// barrier of the NT threads that run diag_body (the whole CTA)
template <int NT>
__device__ __forceinline__ void diag_sync() { asm volatile("bar.sync 4, %0;" ::"n"(NT) : "memory"); }
// T and W tiles + 1/diag.  The 32x32 product scratch of the W assembly lives in the (otherwise unused) upper-right
// quarter of the T tile: rows 0..31, columns 32..63.
static constexpr int SMEM_DIAG = (2 * NB * DT + NB) * 8;

// One 8x8 output tile on the FP64 tensor pipe: C = scale * A[0:8, kb:ke] B[kb:ke, 0:8], A row-major (stride sa),
// B row-major k x n (stride sb), all in shared memory; kb, ke multiples of 4.  Whole warp.
__device__ __forceinline__ void mm_tile8(const double* A, int sa, const double* B, int sb, int kb, int ke, double* C, int sc,
                                         double scale, int lane) {
    const int g = lane >> 2, t = lane & 3;
    double c0 = 0.0, c1 = 0.0;
    // at most 8 k-steps (K <= 32): operands are fetched up front (zeros outside [kb, ke)), then 8 unconditional DMMAs
    double a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k0 = kb + 4 * i;
        const bool v = k0 < ke;
        a[i] = v ? A[g * sa + k0 + t] : 0.0;
        b[i] = v ? B[(k0 + t) * sb + g] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) dmma884(c0, c1, a[i], b[i]);
    C[g * sc + 2 * t] = scale * c0;
    C[g * sc + 2 * t + 1] = scale * c1;
}

// `w_ready()` runs (all threads) once the W image is in global memory, before the L / W^T tile is written back:
// the diagonal CTAs of a chain-bound step raise their flag there -- the waiting TRSMs need W only.
// NT = threads of the calling CTA (256, or 128 in the 64-row panel kernel); all of them must call.
template <int NT, class WReady>
__device__ __forceinline__ void diag_body(const BlockDesc& bd, int blk, int k, const double* __restrict__ sigma,
                                          double* __restrict__ Lbuf, double* __restrict__ wbuf, double ridge,
                                          int32_t* __restrict__ status, double* smem, WReady&& w_ready) {
    static_assert(NT == 128 || NT == 256, "diag_body: 4 or 8 warps");
    constexpr int NWARP = NT / 32;
    double* T = smem;                        // [64][DT] tile, becomes L (lower)
    double* Wf = T + NB * DT;                // [64][DT] W = L^-1 (lower)
    double* dinv_s = Wf + NB * DT;           // [64] 1 / L_ii
    double* Pm = T + 32;                     // [32][DT] product scratch of the W assembly (upper-right quarter of T)
    constexpr int DP = DT;
    const int pc0 = k * NB;
    const int wk = min(NB, bd.mp - pc0);
    const int ld = bd.ld;
    double* Lb = Lbuf + bd.moff;
    const double* src = (k == 0 ? sigma : Lbuf) + bd.moff;     // step 0 reads Sigma, later steps the accumulated tile
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    DIAG_STAMP(0);

    {
        // 128-bit loads of the lower triangle (pairs of columns), all of a thread's loads issued back to back, then consumed
        constexpr int PER = NB * NB / 2 / NT;           // 8 or 16 pairs per thread
        {
            constexpr int u0 = 0;
            double2 tv[PER];
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int idx = tid + (u0 + u) * NT;
                const int a = idx >> 5, b = (idx & 31) * 2;
                const bool ld_it = (a < wk) && (b <= a);
                tv[u] = ld_it ? __ldcg(reinterpret_cast<const double2*>(src + (size_t)(pc0 + a) * ld + pc0 + b)) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < PER; ++u) {
                const int idx = tid + (u0 + u) * NT;
                const int a = idx >> 5, b = (idx & 31) * 2;
                double vx = tv[u].x, vy = (b + 1 <= a) ? tv[u].y : 0.0;      // (a, a + 1) lies above the diagonal
                if (a == b) {
                    if (a >= wk) vx = 1.0;                                  // identity padding
                    else if (k == 0 && pc0 + a < bd.ms) vx += ridge;
                }
                if (a == b + 1) {
                    if (a >= wk) vy = 1.0;
                    else if (k == 0 && pc0 + a < bd.ms) vy += ridge;
                }
                T[a * DT + b] = vx;
                T[a * DT + b + 1] = vy;
                Wf[a * DT + b] = 0.0;
                Wf[a * DT + b + 1] = 0.0;
            }
        }
    }
    diag_sync<NT>();

    bool bad = false;
    DIAG_STAMP(1);
#pragma unroll 1
    for (int jb = 0; jb < 4; ++jb) {
        const int o = 16 * jb;
        if (warp == 0) {
            // ---- 16x16 potrf in registers; lane r (and r+16) holds row r
            const int r = lane & 15;
            double row[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) row[c] = T[(o + r) * DT + o + c];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const double d = __shfl_sync(0xffffffffu, row[j], j);
                if (!(d > 0.0)) bad = true;
                const double inv = rsqrt(d);
                const double sq = d * inv;
                if (lane == j) dinv_s[o + j] = inv;
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
        }
        diag_sync<NT>();                               // [A] L16 and 1/diag are in shared memory
        DIAG_STAMP(2 + 5 * jb);
        const int nrem = 48 - o;                           // rows below this sub-block
        if (warp == 1) {
            // ---- inverse of the 16x16 factor (off the critical path): lane c owns column c of W16;
            // w[i] = -(sum_{k<i} l_ik w[k]) / l_ii
            const int c = lane & 15;
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
                const double di = dinv_s[o + i];
                double wi = -(acc0 + acc1) * di;
                if (i == c) wi = di;
                if (i < c) wi = 0.0;
                w[i] = wi;
            }
            if (lane < 16) {
#pragma unroll
                for (int i = 0; i < 16; ++i) Wf[(o + i) * DT + o + c] = w[i];
            }
        } else if (warp >= 2 && tid - 64 < nrem) {
            // ---- sub-panel below: one row per thread, X L16^T = T_sub by forward substitution
            // (right-looking: each solved x_k is folded into the remaining entries at once, so the chain is
            //  one multiply + one FMA per column)
            const int i = o + 16 + (tid - 64);
            double x[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) x[c] = T[i * DT + o + c];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                x[c] *= dinv_s[o + c];
#pragma unroll
                for (int c2 = c + 1; c2 < 16; ++c2) x[c2] -= x[c] * T[(o + c2) * DT + o + c];
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) T[i * DT + o + c] = x[c];
        }
        diag_sync<NT>();                               // [B] sub-panel solved
        DIAG_STAMP(3 + 5 * jb);
        if (jb == 3) break;
        // ---- next 16x16 diagonal sub-block first (136 packed entries over the threads of warps 1..); then warp 0
        // factors it while the other warps update the rest of the trailing matrix in its shadow (the next barrier [A]
        // publishes that part)
        if (tid >= 32) {
            for (int e = tid - 32; e < 136; e += NT - 32) {
                // entry e of the packed lower triangle -> (i, c)
                int ri = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
                while ((ri + 1) * (ri + 2) / 2 <= e) ++ri;
                while (ri * (ri + 1) / 2 > e) --ri;
                const int ci = e - ri * (ri + 1) / 2;
                const double* xi = T + (o + 16 + ri) * DT + o;
                const double* xc = T + (o + 16 + ci) * DT + o;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                for (int kk = 0; kk < 16; kk += 4) {
                    a0 += xi[kk] * xc[kk]; a1 += xi[kk + 1] * xc[kk + 1];
                    a2 += xi[kk + 2] * xc[kk + 2]; a3 += xi[kk + 3] * xc[kk + 3];
                }
                T[(o + 16 + ri) * DT + o + 16 + ci] -= (a0 + a1) + (a2 + a3);
            }
        }
        diag_sync<NT>();                               // [C] next diagonal sub-block updated
        if (warp != 0) {
            const int h = 32 - o;                          // rows o+32.., columns o+16..row
            for (int idx = tid - 32; idx < h * 64; idx += NT - 32) {
                const int i = o + 32 + (idx >> 6), c = o + 16 + (idx & 63);
                if (c <= i) {
                    const double* xi = T + i * DT + o;
                    const double* xc = T + c * DT + o;
                    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                    for (int kk = 0; kk < 16; kk += 4) {
                        a0 += xi[kk] * xc[kk]; a1 += xi[kk + 1] * xc[kk + 1];
                        a2 += xi[kk + 2] * xc[kk + 2]; a3 += xi[kk + 3] * xc[kk + 3];
                    }
                    T[i * DT + c] -= (a0 + a1) + (a2 + a3);
                }
            }
        }
        DIAG_STAMP(4 + 5 * jb);
    }
    if (warp == 0 && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(&status[blk], 1);
    // ---- W = L^-1 from the four 16x16 inverses, two levels of [[A,0],[B,C]]^-1 = [[A^-1,0],[-C^-1 B A^-1, C^-1]];
    // the small products run on the FP64 tensor pipe, 8x8 output tiles dealt out to the warps.  (The last barrier [B]
    // ordered every write of the loop before this point; the scratch Pm overlays entries of T that hold zeros and are
    // never read again.)
    {
        // level 1: both 32x32 diagonal blocks at once (8 tiles).  P = B A^-1 (16x16 each), then W21 = -C^-1 P
#pragma unroll
        for (int tl = warp; tl < 8; tl += NWARP) {
            const int half = tl >> 2, ob = 32 * half;
            const int r0 = 8 * ((tl >> 1) & 1), c0 = 8 * (tl & 1);
            // A^-1 is lower triangular: only k >= c0 contributes
            mm_tile8(T + (ob + 16 + r0) * DT + ob, DT, Wf + ob * DT + ob + c0, DT, c0, 16, Pm + (16 * half + r0) * DP + c0, DP, 1.0, lane);
        }
        diag_sync<NT>();
#pragma unroll
        for (int tl = warp; tl < 8; tl += NWARP) {
            const int half = tl >> 2, ob = 32 * half;
            const int r0 = 8 * ((tl >> 1) & 1), c0 = 8 * (tl & 1);
            // C^-1 is lower triangular: only k <= r0 + 7 contributes
            mm_tile8(Wf + (ob + 16 + r0) * DT + ob + 16, DT, Pm + (16 * half) * DP + c0, DP, 0, r0 + 8, Wf + (ob + 16 + r0) * DT + ob + c0, DT, -1.0, lane);
        }
        diag_sync<NT>();
        // level 2 (16 tiles): P = L21 W11 (32x32), then W21 = -W22 P
#pragma unroll
        for (int tl = warp; tl < 16; tl += NWARP) {
            const int R0 = 8 * (tl >> 2), C0 = 8 * (tl & 3);
            mm_tile8(T + (32 + R0) * DT, DT, Wf + C0, DT, C0, 32, Pm + R0 * DP + C0, DP, 1.0, lane);
        }
        diag_sync<NT>();
#pragma unroll
        for (int tl = warp; tl < 16; tl += NWARP) {
            const int R0 = 8 * (tl >> 2), C0 = 8 * (tl & 3);
            mm_tile8(Wf + (32 + R0) * DT + 32, DT, Pm + C0, DP, 0, R0 + 8, Wf + (32 + R0) * DT + C0, DT, -1.0, lane);
        }
        diag_sync<NT>();
    }
    DIAG_STAMP(22);
    // ---- write back: W_kk as the shared-memory image the panel kernel of this step copies in with one bulk copy
    // (w_img_off; 16x16 blocks above the diagonal do not exist in it) ...
    double* wb = wbuf + (size_t)blk * (NB * NB);
    for (int i2 = tid; i2 < W_IMG_DOUBLES / 2; i2 += NT) {
        // image position (128-bit unit i2) -> element pair (a, b), (a, b + 1): the inverse of w_img_off
        const int blk16 = i2 >> 7, rb = (blk16 >= 6) ? 3 : (blk16 >= 3) ? 2 : (blk16 >= 1) ? 1 : 0, cb = blk16 - rb * (rb + 1) / 2;
        const int slot = (i2 >> 3) & 7, g8 = ((slot & 3) << 1) | (slot >> 2);
        const int a = 16 * rb + 8 * ((i2 >> 6) & 1) + g8, b = 16 * cb + 2 * ((i2 & 7) ^ slot);
        reinterpret_cast<double2*>(wb)[i2] = make_double2((b <= a) ? Wf[a * DT + b] : 0.0, (b + 1 <= a) ? Wf[a * DT + b + 1] : 0.0);
    }
    w_ready();
    // ... and the tile itself: lower = L_kk, strict upper = W_kk^T (used by the back substitution); wk is a multiple of 8
    for (int idx = tid; idx < NB * NB / 2; idx += NT) {
        const int a = idx >> 5, b = (idx & 31) * 2;
        if (a < wk && b < wk)
            *reinterpret_cast<double2*>(Lb + (size_t)(pc0 + a) * ld + pc0 + b) =
                make_double2((b <= a) ? T[a * DT + b] : Wf[b * DT + a], (b + 1 <= a) ? T[a * DT + b + 1] : Wf[(b + 1) * DT + a]);
    }
    DIAG_STAMP(23);
}

__global__ void __launch_bounds__(CHOL_THREADS, 3)
chol_diag_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ items, int32_t k,
                 const double* __restrict__ sigma, double* __restrict__ Lbuf, double* __restrict__ wbuf,
                 double ridge, int32_t* __restrict__ status) {
    extern __shared__ __align__(16) double smem[];
    const int blk = items[blockIdx.x];
    const BlockDesc bd = blocks[blk];
    diag_body<CHOL_THREADS>(bd, blk, k, sigma, Lbuf, wbuf, ridge, status, smem, []() {});
}

// Look-ahead SYRK of one 16-row slot WL of a 64-row half: acc2 (rows 16WL.., columns 0 .. 16WL+15, lower triangle
// of the tile) += A A^T with both operands read from the half tile in shared memory (stride LDW).
template <int WL>
__device__ __forceinline__ void syrk_rows(const double* __restrict__ A, const double* __restrict__ B,
                                          double (&acc2)[2][8][2]) {
#pragma unroll 2
    for (int s8 = 0; s8 < NB / 8; ++s8) {
        const double2 a0 = *reinterpret_cast<const double2*>(A + s8 * 8);
        const double2 a1 = *reinterpret_cast<const double2*>(A + 8 * LDW + s8 * 8);
#pragma unroll
        for (int c = 0; c <= 2 * WL + 1; ++c) {
            const double2 b = *reinterpret_cast<const double2*>(B + c * 8 * LDW + s8 * 8);
            if (c <= 2 * WL) dmma884(acc2[0][c][0], acc2[0][c][1], a0.x, b.x);   // rows 16WL..+7 end at column block 2WL
            dmma884(acc2[1][c][0], acc2[1][c][1], a1.x, b.x);
            if (c <= 2 * WL) dmma884(acc2[0][c][0], acc2[0][c][1], a0.y, b.y);
            dmma884(acc2[1][c][0], acc2[1][c][1], a1.y, b.y);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Rows below the diagonal tile of panel k:
//   acc = L_i,0:k L_k,0:k^T - K_ik = -C (DMMA, cp.async ring), -L_ik = acc W_kk^T (DMMA, TRSM as a GEMM whose A
//   operand is the accumulator fragment itself: with the K order permuted so that lane t owns k = 2t, 2t+1 of every
//   8-wide group, the C fragment of m8n8k4 IS its A fragment -- no shared-memory round trip), then the look-ahead:
//   each of the (up to two) 64-row halves of the macro tile immediately applies T_ii -= L_ik L_ik^T to ITS OWN
//   diagonal tile (one writer per tile and step, so no atomics).  With kFuseDiag the CTA of macro tile 0 -- whose
//   first half is the diagonal tile of panel k+1, now fully updated -- also factors and inverts that tile, so a
//   panel step is ONE launch and the next step's W is ready when it starts.
// ------------------------------------------------------------------------------------------
// item: x = block, y = macro tile, z = slice | nslices << 8, w = split group id.
// fuse_end: the CTA of macro tile 0 factors the diagonal tile of panel k+1 when it is done.  wait_w: W_k is being
// produced by a diagonal CTA of THIS launch (see chol_panel_kernel): the main loop runs without it and the TRSM
// waits for dflag[block] >= k+1.
__device__ __forceinline__ void panel_body(const BlockDesc& bd, const int4 item, int k, const double* __restrict__ sigma,
                                           double* __restrict__ Lbuf, double* __restrict__ wbuf, int64_t wpar,
                                           int64_t wpar_next, double ridge, double* __restrict__ scratch,
                                           int32_t* __restrict__ counters, int32_t group_base,
                                           int32_t* __restrict__ status, const int32_t* __restrict__ dflag, bool fuse_end,
                                           bool wait_w, double* smem, int* s_last_p) {
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
    // 16-row slot of this warp.  The second 64-row half is mirrored: warps w and w+4 share an SM sub-partition, and
    // the look-ahead SYRK of slot wl costs ~(wl+1) units, so (wl, 3-wl) pairs balance the four tensor pipes.
    const int grp = warp >> 2;
    const int wl = grp ? 3 - (warp & 3) : (warp & 3);
    const int wrow = 4 * grp + wl;

    const int slice = item.z & 0xFF, nsl = item.z >> 8;
    // accumulators start at -K_ik (slice 0 only); the preload is issued after the ring prologue, so the Sigma tile's
    // HBM latency and the first stages' latency overlap: after the loop acc = L_i,0:k L_k,0:k^T - K_ik = -C
    double acc[2][8][2];
    // split-K: slice s of nsl owns 64-wide K blocks [k*s/nsl, k*(s+1)/nsl)
    const int kb = (k * slice) / nsl * NB, ke = (k * (slice + 1)) / nsl * NB;
    const double* wsrc = wbuf + wpar + (size_t)item.x * (NB * NB);
    gemm_nt_core(Lb + (size_t)r0 * ld, Lb + (size_t)pc0 * ld, ld, prow, wk, kb, ke,
                 wait_w ? nullptr : wsrc, wrow, smem, acc, [&]() {
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int rl = 16 * wrow + 8 * f + g, cc = 8 * c + 2 * t;
                double2 a = make_double2(0.0, 0.0);
                if (slice == 0 && rl < prow && cc < wk) a = *reinterpret_cast<const double2*>(Sb + (size_t)(r0 + rl) * ld + pc0 + cc);
                acc[f][c][0] = -a.x;
                acc[f][c][1] = -a.y;
            }
    });
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
        if (!s_last) return;
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

    // Shared memory now: stage s_w holds W_kk ([64][LDW], W[c][c'] = (L_kk^-1)[c][c'], zero above the diagonal);
    // the other two stages take the two 64-row halves of -L_ik ([64][LDW] each).
    const int s_w = ((ke - kb) / KC) % NST;
    if (wait_w) {
        // W_k comes from a diagonal CTA of this launch (lower blockIdx, so it is resident or done): acquire, then load
        if (tid == 0) {
            const volatile int32_t* f = dflag + item.x;
            uint32_t spins = 0;                 // bounded like the TMA kernel's wait (see there)
            while (*f < k + 1) {
                __nanosleep(40);
                if (++spins > (1u << 27)) { atomicOr(&status[item.x], 4); break; }
            }
            __threadfence();
        }
        __syncthreads();
        double* Ws = smem + s_w * STAGE;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int idx = tid + u * CHOL_THREADS;
            const int row = idx >> 5, ch = (idx & 31) * 2;
            const bool v = (ch >> 4) <= (row >> 4);
            cp_async16(Ws + row * LDW + ch, wsrc + (v ? w_img_off(row, ch) : 0), v);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
    }
    const double* Wsm = smem + s_w * STAGE;
    double* Lh = smem + ((s_w + 1 + grp) % NST) * STAGE;
    const bool active = (16 * wrow < prow);
    const int trow0 = r0 + 64 * grp;              // first global row of this warp group's 64-row half
    const int tw = min(NB, bd.mp - trow0);        // extent of its diagonal tile (<= 0: z / test rows, no tile)
    const bool glook = (trow0 < bd.mp && wk == NB);   // group-uniform (a narrow last panel has no diagonal tile below)
    const bool look = glook && (16 * wl < tw);

    // ---- TRSM, 32 output columns at a time (keeps accumulators + outputs within the register budget)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        double out[2][4][2];
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int cq = 0; cq < 4; ++cq) out[f][cq][0] = out[f][cq][1] = 0.0;
        if (active) {
#pragma unroll
            for (int cp = 0; cp < 4 * h + 4; ++cp) {
#pragma unroll
                for (int cq = 0; cq < 4; ++cq) {
                    const int c = 4 * h + cq;
                    if (c >= cp) {                                  // W is lower triangular (static: no predicated mma)
                        const double2 b = *reinterpret_cast<const double2*>(Wsm + (8 * c + g) * LDW + 8 * cp + 2 * t);
                        dmma884(out[0][cq][0], out[0][cq][1], acc[0][cp][0], b.x);
                        dmma884(out[1][cq][0], out[1][cq][1], acc[1][cp][0], b.x);
                        dmma884(out[0][cq][0], out[0][cq][1], acc[0][cp][1], b.y);
                        dmma884(out[1][cq][0], out[1][cq][1], acc[1][cp][1], b.y);
                    }
                }
            }
        }
        // out = -L_ik: L to global, -L into this group's half tile (operand of the look-ahead SYRK; the sign cancels)
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int cq = 0; cq < 4; ++cq) {
                const int rl = 16 * wrow + 8 * f + g, cc = 8 * (4 * h + cq) + 2 * t;
                if (rl < prow && cc < wk)
                    *reinterpret_cast<double2*>(Lb + (size_t)(r0 + rl) * ld + pc0 + cc) =
                        make_double2(-out[f][cq][0], -out[f][cq][1]);
                *reinterpret_cast<double2*>(Lh + (16 * wl + 8 * f + g) * LDW + cc) = make_double2(out[f][cq][0], out[f][cq][1]);
            }
    }

    // ---- look-ahead: T_ii -= L_ik L_ik^T on this group's diagonal tile
    if (glook) {
        // the old tile values are loaded straight into the accumulators and NOT touched before the barrier (their
        // latency overlaps it); sign and ridge are applied after it
        double acc2[2][8][2];
        const double* src = (k == 0 ? sigma : Lbuf) + bd.moff;    // first touch reads Sigma (+ ridge)
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int rl = 16 * wl + 8 * f + g, cc = 8 * c + 2 * t;   // within the tile
                double2 v = make_double2(0.0, 0.0);
                if (look && c <= 2 * wl + 1 && rl < tw && cc <= rl)
                    v = __ldcg(reinterpret_cast<const double2*>(src + (size_t)(trow0 + rl) * ld + trow0 + cc));
                acc2[f][c][0] = v.x;
                acc2[f][c][1] = v.y;
            }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");      // the four warps of this half
        if (look) {
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int rl = 16 * wl + 8 * f + g, cc = 8 * c + 2 * t;
                    double vx = acc2[f][c][0], vy = acc2[f][c][1];
                    if (k == 0 && trow0 + rl < bd.ms) {
                        if (cc == rl) vx += ridge;
                        if (cc + 1 == rl) vy += ridge;                    // odd rows: the diagonal is the pair's second element
                    }
                    acc2[f][c][0] = -vx;
                    acc2[f][c][1] = -vy;
                }
            const double* A = Lh + (16 * wl + g) * LDW + 2 * t;
            const double* B = Lh + g * LDW + 2 * t;
            switch (wl) {                                              // static column count per row slot: no predicated mma
                case 0: syrk_rows<0>(A, B, acc2); break;
                case 1: syrk_rows<1>(A, B, acc2); break;
                case 2: syrk_rows<2>(A, B, acc2); break;
                default: syrk_rows<3>(A, B, acc2); break;
            }
            // acc2 = L L^T - T_old  =>  T_new = -acc2
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int rl = 16 * wl + 8 * f + g, cc = 8 * c + 2 * t;   // within the tile
                    if (c <= 2 * wl + 1 && rl < tw && cc <= rl)
                        *reinterpret_cast<double2*>(Lb + (size_t)(trow0 + rl) * ld + trow0 + cc) =
                            make_double2(-acc2[f][c][0], -acc2[f][c][1]);
                }
        }
    }
    if (fuse_end) {
        // macro tile 0: rows r0 .. r0+63 are the diagonal tile of panel k+1, which this step completed
        if (item.y == 0 && r0 < bd.mp && wk == NB) {
            __syncthreads();        // T_new is in L2 for the whole CTA; shared memory is free again
            diag_body<CHOL_THREADS>(bd, item.x, k + 1, sigma, Lbuf, wbuf + wpar_next, ridge, status, smem, []() {});
        }
    }
}

// One panel step of a batch.  CTAs [0, n_diag_first) factor the diagonal tile of panel k of their block (when step k-1
// deferred it: chain-bound steps, where the factorisation then overlaps this step's main loops instead of extending
// step k-1) and publish it through dflag; the others are panel CTAs.  CTAs are dispatched in blockIdx order, so a
// panel CTA never waits for a diagonal CTA that is not resident or finished.
__global__ void __launch_bounds__(CHOL_THREADS, 2)
chol_panel_kernel(const BlockDesc* __restrict__ blocks, const int4* __restrict__ items,
                  const int32_t* __restrict__ diag_items, int32_t n_diag_first, int32_t k,
                  const double* __restrict__ sigma, double* __restrict__ Lbuf, double* __restrict__ wbuf, int64_t wpar,
                  int64_t wpar_next, double ridge, double* __restrict__ scratch, int32_t* __restrict__ counters,
                  int32_t group_base, int32_t* __restrict__ status, int32_t* __restrict__ dflag, int32_t fuse_end) {
    extern __shared__ __align__(16) double smem[];
    __shared__ int s_last;
    // Programmatic dependent launch (only when the host set the launch attribute; no-ops otherwise): the CTAs of step
    // k+1 become resident while step k is still running -- they take SM slots as lower-priority work frees them --
    // and block here until step k has completed and its writes are visible.  Everything below reads step k's output.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if ((int)blockIdx.x < n_diag_first) {
        const int blk = diag_items[blockIdx.x];
        const BlockDesc bd = blocks[blk];
        diag_body<CHOL_THREADS>(bd, blk, k, sigma, Lbuf, wbuf + wpar, ridge, status, smem, [&]() {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) atomicExch(dflag + blk, k + 1);      // W_k is visible: the step's TRSMs may start
        });
        return;
    }
    const int4 item = items[blockIdx.x - n_diag_first];
    const BlockDesc bd = blocks[item.x];
    panel_body(bd, item, k, sigma, Lbuf, wbuf, wpar, wpar_next, ridge, scratch, counters, group_base, status, dflag,
               fuse_end != 0, n_diag_first > 0, smem, &s_last);
}

// ------------------------------------------------------------------------------------------
// TMA panel kernel: the same panel step as chol_panel_kernel (same arithmetic, same work lists), restructured around
// what the round-1 profile of that kernel showed -- 29 % of its warp samples sat in the per-chunk __syncthreads of the
// cp.async ring, 8 % in cp.async.wait, every CTA paid two dependent descriptor loads, a cold ring and the HBM latency
// of its Sigma tile before its first DMMA:
//   * thread 0 is the PRODUCER: it walks the CTA's items and feeds a 3-stage ring of [64 rows x 16 doubles] boxes
//     (one per 64 macro-tile rows, one for the 64 panel rows) with 4-D tiled TMA loads (cp.async.bulk.tensor,
//     SWIZZLE_128B; one tensor map per block, in the plan blob), completion on `full` mbarriers;
//   * the CONSUMER warps never meet in the K loop: each waits for `full[s]`, reads its fragments, and releases
//     the stage on `empty[s]` -- no CTA-wide barrier, no per-thread address arithmetic, no cp.async bookkeeping;
//   * bank conflicts: the tensor map splits a row index r = 8a + 2b + c into (b, c, a) and lists b before c, so the
//     rows of an 8-row group land in shared memory in the order 0,2,4,6,1,3,5,7; with the 128-byte swizzle the eight
//     lanes of every 128-bit load phase (fragment rows g = 2j, 2j+1, four 16-byte chunks each) then hit eight distinct
//     bank groups -- unpadded stages;
//   * W_kk arrives as a ready-made 20 KB image (w_img_off) by ONE bulk copy while the ring is busy; the epilogue tiles
//     (the 64-row halves of -L_ik) reuse the ring boxes in the same permuted/swizzled format, so one address function
//     serves all;
//   * a CTA may own several consecutive items (macro tiles of the same block and step): the producer is then already
//     waiting with the next item's loads, its Sigma tile has been pulled into L2 during the previous item, and W_kk
//     is loaded once per block;
//   * `pf` > 0: the rows of chunk c + pf are pulled into L2 (tensor prefetch) when chunk c is issued, so that the two
//     chunks the ring keeps in flight come from L2 rather than from HBM.
// Two shapes (template NG = 64-row warp groups per CTA):
//   NG = 2: 128-row macro tiles, 256 threads, 93 KB, two CTAs per SM (128 registers);
//   NG = 1:  64-row macro tiles, 128 threads, 69 KB, THREE CTAs per SM (168 registers, no spills).  The round-2 profile
//            of the NG = 2 shape read: FP64 tensor pipe 66 % busy with each CTA "DMMA-ready" only 42 % of its time --
//            exactly what two independent CTAs per SM give (1 - 0.58^2); what is missing is a third independent
//            instruction stream per SM, not more warps per CTA.  Partially filled macro tiles also idle fewer warps.
// ------------------------------------------------------------------------------------------
#ifndef CHOL_VARIANT
#define CHOL_VARIANT 0          // hook for tuning/debugging variants of this file (tools/build_variants.sh)
#endif
static constexpr int BOXB = NB * KC * 8;                 // bytes of one [64 x 16] box (8 KB)
static constexpr int TP_NST = 3;                         // ring stages
static_assert(KC == 16, "a box row is one 128-byte swizzle span");

template <int NG>
struct TpCfg {
    static constexpr int NT = 128 * NG;                  // threads; thread 0 doubles as the TMA producer
    static constexpr int NW = 4 * NG;                    // warps
    static constexpr int TMR = 64 * NG;                  // rows per macro tile
    static constexpr int SB = NG + 1;                    // boxes per stage: NG row boxes + the panel-row box
    static constexpr int RING = TP_NST * SB * BOXB;      // also the NG L halves of the epilogue (4 boxes each)
    static constexpr int SMEM = RING + W_IMG_BYTES + 1024;   // + W_kk image + alignment slack
    static constexpr int MINB = (NG == 2) ? 2 : 3;       // CTAs per SM
    static_assert(4 * NG <= TP_NST * SB, "the L halves must fit into the ring");
    static_assert(RING + W_IMG_BYTES >= SMEM_DIAG, "the fused diagonal factorisation reuses the panel kernel's shared memory");
};

struct TpBars {
    uint64_t full[TP_NST], empty[TP_NST];
    uint64_t w_full, w_free, ring_free;
};

__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int32_t c0, int32_t c1, int32_t c2, int32_t c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(tmap), "r"(c0), "r"(c1),
                 "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int32_t x, int32_t y) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(x), "r"(y) : "memory");
}

template <int WL>
__device__ __forceinline__ void syrk_rows_pt(uint32_t Lh, uint32_t lrow, uint32_t x0, double (&acc2)[2][8][2]) {
#pragma unroll 2
    for (int s8 = 0; s8 < NB / 8; ++s8) {
        const uint32_t bx = Lh + (uint32_t)(s8 >> 1) * BOXB + lrow + (x0 ^ ((uint32_t)(s8 & 1) << 6));
        const double2 a0 = lds_f64x2(bx + (2 * WL) * 1024);
        const double2 a1 = lds_f64x2(bx + (2 * WL + 1) * 1024);
#pragma unroll
        for (int c = 0; c <= 2 * WL + 1; ++c) {
            const double2 b = lds_f64x2(bx + c * 1024);
            if (c <= 2 * WL) dmma884(acc2[0][c][0], acc2[0][c][1], a0.x, b.x);
            dmma884(acc2[1][c][0], acc2[1][c][1], a1.x, b.x);
            if (c <= 2 * WL) dmma884(acc2[0][c][0], acc2[0][c][1], a0.y, b.y);
            dmma884(acc2[1][c][0], acc2[1][c][1], a1.y, b.y);
        }
    }
}

// geometry of one item (TMR = rows per macro tile)
struct TpGeom {
    int pc0, wk, r0, prow, kb, ke, slice, nsl;
};
template <int TMR>
__device__ __forceinline__ TpGeom tp_geom(const BlockDesc& bd, const int4 item, int k) {
    TpGeom q;
    q.pc0 = k * NB;
    q.wk = min(NB, bd.mp - q.pc0);
    q.r0 = q.pc0 + q.wk + item.y * TMR;
    q.prow = min(TMR, bd.nrows - q.r0);
    q.slice = item.z & 0xFF;
    q.nsl = item.z >> 8;
    q.kb = (k * q.slice) / q.nsl * NB;
    q.ke = (k * (q.slice + 1)) / q.nsl * NB;
    return q;
}

// lmaps: one 4-D tensor map per block over its matrix in Lbuf (dims {ld, 4, 2, rows/8}, box {16, 4, 2, 8}; `perm` = 0:
// plain 2-D maps {ld, rows}, box {16, 64} -- rows in natural order, two-way bank conflicts -- if the driver refuses the
// permuting strides).
// CTA c (after the n_diag_first diagonal CTAs): c < n_single owns item c; the others own `tpc` consecutive items.
template <int NG>
__global__ void __launch_bounds__(TpCfg<NG>::NT, TpCfg<NG>::MINB)
chol_panel_tma_kernel(const BlockDesc* __restrict__ blocks, const int4* __restrict__ items, int32_t n_items,
                      int32_t n_single, int32_t tpc, const int32_t* __restrict__ diag_items, int32_t n_diag_first,
                      int32_t k, const CUtensorMap* __restrict__ lmaps, int32_t perm, int32_t pf,
                      const double* __restrict__ sigma, double* __restrict__ Lbuf,
                      double* __restrict__ wbuf, int64_t wpar, int64_t wpar_next, double ridge,
                      double* __restrict__ scratch, int32_t* __restrict__ counters, int32_t group_base,
                      int32_t* __restrict__ status, int32_t* __restrict__ dflag, int32_t fuse_end) {
    using C = TpCfg<NG>;
    constexpr int NT = C::NT, NW = C::NW, TMR = C::TMR, SB = C::SB;
    extern __shared__ __align__(16) uint8_t tp_smem_raw[];
    __shared__ __align__(8) TpBars bars;
    __shared__ int s_last;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* smem_al = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tp_smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sbase = smem_u32(smem_al);

    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if ((int)blockIdx.x < n_diag_first) {
        const int blk = diag_items[blockIdx.x];
        const BlockDesc bd = blocks[blk];
        diag_body<NT>(bd, blk, k, sigma, Lbuf, wbuf + wpar, ridge, status, reinterpret_cast<double*>(smem_al), [&]() {
            __threadfence();
            diag_sync<NT>();
            if (threadIdx.x == 0) atomicExch(dflag + blk, k + 1);
        });
        return;
    }
    const bool wait_w = n_diag_first > 0;
    const int cta = (int)blockIdx.x - n_diag_first;
    int i0, n_my;
    if (cta < n_single) { i0 = cta; n_my = 1; }
    else { i0 = n_single + (cta - n_single) * tpc; n_my = min(tpc, n_items - i0); }

    if (tid == 0) {
        for (int s = 0; s < TP_NST; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], NW); }
        mbar_init(&bars.w_full, 1);
        mbar_init(&bars.w_free, NW);
        mbar_init(&bars.ring_free, NW);
        mbar_fence_init();
    }
    __syncthreads();

    uint8_t* wimg = smem_al + C::RING;                         // W_kk image (w_img_off)
    const uint32_t wbox = sbase + C::RING;
    const int g = lane >> 2, t = lane & 3;
    const int sigp = (g >> 1) | ((g & 1) << 2);                 // row slot of fragment row g in a row-permuted 8-row group
    const int sig = perm ? sigp : g;                            // ... in the ring boxes (natural order with plain 2-D maps)
    const uint32_t lrow = (uint32_t)sig * 128u;
    const uint32_t x0 = (uint32_t)((t ^ sig) & 7) << 4;          // 16-byte chunk of columns 2t, 2t+1 of the first 8-column group
    const uint32_t wlrow = (uint32_t)sigp * 128u;                // the same two for the W image (always permuted)
    const uint32_t wx0 = (uint32_t)((t ^ sigp) & 7) << 4;
    const int grp = warp >> 2;
    // Ring bookkeeping: chunk c of EVERY item uses stage (c + 2) % 3, so chunk 0 always lands in stage 2 -- the one stage the
    // epilogue (whose L halves live in boxes 0..3 = stages 0 and 1 of the 64-row shape) leaves alone, which lets thread 0
    // issue the NEXT item's first chunk while this item's TRSM / SYRK run (`pre`).  Parity bit s of pprod / pcons = uses of
    // stage s issued by the producer / consumed by this thread so far.
    uint32_t pprod = 0, pcons = 0, wcount = 0;                   // (wcount: W loads so far; all threads keep count)
    const int pf_dist = pf & 0xFF;
    const bool next_pf = (NG == 1) && ((pf >> 8) & 1);
    bool pre = false;
    int prev_blk = -1;

    for (int j = 0; j < n_my; ++j) {
        const int4 item = items[i0 + j];
        const BlockDesc bd = blocks[item.x];
        const TpGeom q = tp_geom<TMR>(bd, item, k);
        const int pc0 = q.pc0, wk = q.wk, r0 = q.r0, prow = q.prow, ld = bd.ld;
        double* Lb = Lbuf + bd.moff;
        const double* Sb = sigma + bd.moff;
        // 16-row slot of this warp.  The look-ahead SYRK of slot wl costs ~(wl + 1) units and warp w always runs on SM
        // sub-partition w & 3, so the slots are mirrored -- between the two groups of a 256-thread CTA, between odd and even
        // items of the 128-thread shape -- to spread the heavy slots over the four tensor pipes.  (A function of the ITEM:
        // the split-K slices of one item, summed thread by thread, must all use the same mapping.)
        const bool mirror = (NG == 2) ? (grp != 0) : (((item.x + item.y) & 1) != 0);
        const int wl = mirror ? 3 - (warp & 3) : (warp & 3);
        const int wrow = 4 * grp + wl;
        const bool active = (16 * wrow < prow);
        const bool load_w = (item.x != prev_blk);
        prev_blk = item.x;
        const int nchunk = (q.ke - q.kb) / KC;
        const CUtensorMap* lm = lmaps + item.x;
        const bool two_row_boxes = (NG == 2) && prow > NB;

        // ---- producer duties of thread 0.  Chunk c goes to stage (c + 2) % 3, which may be written once all consumer
        // warps have released the stage's previous use.  (Geometry passed in: the same code issues the next item's chunk 0.)
        auto issue_chunk_of = [&](const CUtensorMap* lm_, int kb_, int r0_, int pc0_, bool two_, int nchunk_, int c) {
            const int s = (c + 2) % TP_NST;
            mbar_wait(&bars.empty[s], ((pprod >> s) & 1u) ^ 1u);
            pprod ^= 1u << s;
            mbar_expect_tx(&bars.full[s], (uint32_t)((two_ ? 3 : 2) * BOXB));
            const int k0 = kb_ + c * KC;
            if (perm) {
                const uint32_t st = sbase + (uint32_t)s * SB * BOXB;
                tma_load_4d(st, lm_, k0, 0, 0, r0_ >> 3, &bars.full[s]);
                if (two_) tma_load_4d(st + BOXB, lm_, k0, 0, 0, (r0_ >> 3) + 8, &bars.full[s]);
                tma_load_4d(st + NG * BOXB, lm_, k0, 0, 0, pc0_ >> 3, &bars.full[s]);
                if (pf_dist > 0 && c + pf_dist < nchunk_) {
                    tma_prefetch_4d(lm_, k0 + pf_dist * KC, 0, 0, r0_ >> 3);
                    if (two_) tma_prefetch_4d(lm_, k0 + pf_dist * KC, 0, 0, (r0_ >> 3) + 8);
                }
            } else {
                uint8_t* sp = smem_al + (size_t)s * SB * BOXB;
                tma_load_2d(sp, lm_, k0, r0_, &bars.full[s]);
                if (two_) tma_load_2d(sp + BOXB, lm_, k0, r0_ + NB, &bars.full[s]);
                tma_load_2d(sp + NG * BOXB, lm_, k0, pc0_, &bars.full[s]);
                if (pf_dist > 0 && c + pf_dist < nchunk_) {
                    tma_prefetch_2d(lm_, k0 + pf_dist * KC, r0_);
                    if (two_) tma_prefetch_2d(lm_, k0 + pf_dist * KC, r0_ + NB);
                }
            }
        };
        auto issue_chunk = [&](int c) { issue_chunk_of(lm, q.kb, r0, pc0, two_row_boxes, nchunk, c); };
        auto issue_w = [&]() {
            mbar_expect_tx(&bars.w_full, W_IMG_BYTES);
            bulk_g2s(wimg, wbuf + wpar + (size_t)item.x * (NB * NB), W_IMG_BYTES, &bars.w_full);
        };
        if (tid == 0) {
            if (load_w) tensormap_acquire(lm);
            // the ring first (it gates the main loop), then W_kk (needed only by the epilogue)
            if (j > 0) mbar_wait(&bars.ring_free, (uint32_t)((j - 1) & 1));
            if (nchunk > 0 && !pre) issue_chunk(0);          // (pre: already in flight since the previous item's epilogue)
            if (nchunk > 1) issue_chunk(1);
            pre = false;
            if (load_w && !wait_w) {
                if (j > 0) mbar_wait(&bars.w_free, (uint32_t)((j - 1) & 1));
                issue_w();
            }
        }

        // accumulators start at -K_ik (slice 0 only): after the loop acc = L_i,0:k L_k,0:k^T - K_ik = -C
        double acc[2][8][2];
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int rl = 16 * wrow + 8 * f + g, cc = 8 * c + 2 * t;
                double2 a = make_double2(0.0, 0.0);
                if (q.slice == 0 && rl < prow && cc < wk) a = *reinterpret_cast<const double2*>(Sb + (size_t)(r0 + rl) * ld + pc0 + cc);
                acc[f][c][0] = -a.x;
                acc[f][c][1] = -a.y;
            }

        // ---- main loop: no CTA-wide barrier; thread 0 keeps the ring two chunks ahead.
        // A stage is handed back one iteration late, AFTER the wait for the next chunk: by then every DMMA that consumes
        // its fragments has been issued (in-order issue), and an issued DMMA has its operands, i.e. the shared-memory loads
        // have really returned.  Until round 2 the release sat right after the last LDS of the stage -- ptxas is free to
        // move an mbarrier arrive above DMMAs (it did), an arrive does not wait for LDS that are still queued in the
        // shared-memory pipe, and under load the next TMA box then landed beneath them: run-to-run differences of ~1e-5
        // in a few percent of the blocks once three CTAs per SM kept that pipe busy (tools/stream_vs_resident.py,
        // tools/l_diff.py: one warp's 16 rows, last column groups of a panel).  The spin loop of the wait is a point the
        // arrive cannot be hoisted across.
        int s = TP_NST - 1, s_prev = 0;
        for (int kc = 0; kc < nchunk; ++kc) {
            mbar_wait(&bars.full[s], (pcons >> s) & 1u);
            pcons ^= 1u << s;
            if (kc > 0) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars.empty[s_prev]);
            }
            if (tid == 0 && kc + 2 < nchunk) issue_chunk(kc + 2);
            if (active) {
                const uint32_t st = sbase + (uint32_t)s * SB * BOXB;
                const uint32_t Pa = st + (uint32_t)grp * BOXB + (uint32_t)(2 * wl) * 1024 + lrow;
                const uint32_t Qa = st + NG * BOXB + lrow;
#pragma unroll
                for (int s8 = 0; s8 < 2; ++s8) {
                    const uint32_t xo = x0 ^ ((uint32_t)s8 << 6);
                    const double2 a0 = lds_f64x2(Pa + xo);
                    const double2 a1 = lds_f64x2(Pa + 1024 + xo);
#pragma unroll
                    for (int hc = 0; hc < 2; ++hc) {
                        double2 b[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) b[c] = lds_f64x2(Qa + (uint32_t)(4 * hc + c) * 1024 + xo);
                        // eight independent accumulators per pass (x, then y): dependent DMMAs are 8 apart
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            dmma884(acc[0][4 * hc + c][0], acc[0][4 * hc + c][1], a0.x, b[c].x);
                            dmma884(acc[1][4 * hc + c][0], acc[1][4 * hc + c][1], a1.x, b[c].x);
                        }
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            dmma884(acc[0][4 * hc + c][0], acc[0][4 * hc + c][1], a0.y, b[c].y);
                            dmma884(acc[1][4 * hc + c][0], acc[1][4 * hc + c][1], a1.y, b[c].y);
                        }
                    }
                }
            }
            s_prev = s;
            s = (s == TP_NST - 1) ? 0 : s + 1;
        }
        if (wait_w && load_w && tid == 0) {
            // W_k is being produced by a diagonal CTA of THIS launch (lower blockIdx: resident or done)
            const volatile int32_t* fl = dflag + item.x;
            // bounded: the diagonal CTAs have lower block indices and are dispatched first, so the flag comes within tens of
            // microseconds; if it has not come after ~5 s something is broken, and a flagged block beats a hung device
            uint32_t spins = 0;
            while (*fl < k + 1) {
                __nanosleep(40);
                if (++spins > (1u << 27)) { atomicOr(&status[item.x], 4); break; }      // DBSLMM_B200_BLK_SYNC_TIMEOUT
            }
            __threadfence();
            fence_proxy_async();
            issue_w();
        }
        __syncthreads();          // every warp is done with the ring: its boxes become the epilogue's L halves
        if (nchunk > 0) {         // the last chunk's stage (behind the barrier: see the main loop)
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars.empty[s_prev]);
        }

        {
            // this item's diagonal tiles (old values of the look-ahead update, read after the TRSM) -> L2 now
            const int hrow0 = r0 + 64 * (tid >> 6), hr = tid & 63;
            if (tid < 64 * NG && wk == NB && hrow0 + hr < bd.mp)
                l2_prefetch_bulk((k == 0 ? sigma : Lbuf) + bd.moff + (size_t)(hrow0 + hr) * ld + hrow0,
                                 (uint32_t)(min(NB, bd.mp - hrow0) * 8));
        }
        if (j + 1 < n_my) {
            // the next item's Sigma tile (its accumulator preload) -> L2 while this item's epilogue runs
            const int4 item2 = items[i0 + j + 1];
            const BlockDesc bd2 = blocks[item2.x];
            const TpGeom q2 = tp_geom<TMR>(bd2, item2, k);
            if (q2.slice == 0 && tid < q2.prow)
                l2_prefetch_bulk(sigma + bd2.moff + (size_t)(q2.r0 + tid) * bd2.ld + q2.pc0, (uint32_t)(q2.wk * 8));
            if (next_pf && tid == 0 && q2.ke > q2.kb) {
                // ... and its first chunk into stage 2, which the epilogue does not touch: the next main loop starts warm
                const CUtensorMap* lm2 = lmaps + item2.x;
                if (item2.x != item.x) tensormap_acquire(lm2);
                issue_chunk_of(lm2, q2.kb, q2.r0, q2.pc0, false, (q2.ke - q2.kb) / KC, 0);
                pre = true;
            }
        }

        if (q.nsl > 1) {
            // split-K (single-item CTAs only): partial sums to scratch, the last arriver adds them up IN SLICE ORDER
            double* part = scratch + ((size_t)(item.w - group_base) * q.nsl) * (TMR * NB);
            double2* mine = reinterpret_cast<double2*>(part + (size_t)q.slice * (TMR * NB)) + tid;
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int c = 0; c < 8; ++c) mine[(f * 8 + c) * NT] = make_double2(acc[f][c][0], acc[f][c][1]);
            __threadfence();
            __syncthreads();
            if (tid == 0) s_last = (atomicAdd(&counters[item.w], 1) == q.nsl - 1);
            __syncthreads();
            if (!s_last) {
                if (load_w) mbar_wait(&bars.w_full, wcount & 1u);     // no bulk copy may be in flight when the CTA exits
                return;
            }
            __threadfence();
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[f][c][0] = acc[f][c][1] = 0.0;
            for (int sl = 0; sl < q.nsl; ++sl) {
                const double2* src2 = reinterpret_cast<const double2*>(part + (size_t)sl * (TMR * NB)) + tid;
#pragma unroll
                for (int f = 0; f < 2; ++f)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const double2 v = __ldcg(src2 + (f * 8 + c) * NT);
                        acc[f][c][0] += v.x;
                        acc[f][c][1] += v.y;
                    }
            }
        }

        if (load_w) { mbar_wait(&bars.w_full, wcount & 1u); ++wcount; }
        const uint32_t Lh = sbase + (uint32_t)grp * 4 * BOXB;          // this warp group's half of -L_ik
        const int trow0 = r0 + 64 * grp;
        const int tw = min(NB, bd.mp - trow0);
        const bool glook = (trow0 < bd.mp && wk == NB);
        const bool look = glook && (16 * wl < tw);

        // ---- TRSM: -L_ik = (-C) W_kk^T, A operand = the accumulator fragments, 32 output columns at a time
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double out[2][4][2];
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int cq = 0; cq < 4; ++cq) out[f][cq][0] = out[f][cq][1] = 0.0;
            if (active) {
#pragma unroll
                for (int cp = 0; cp < 4 * h + 4; ++cp) {
#pragma unroll
                    for (int cq = 0; cq < 4; ++cq) {
                        const int c = 4 * h + cq;
                        if (c >= cp) {
                            // W rows 8c.., columns 8cp..: 16x16 block (c >> 1, cp >> 1) of the image
                            const double2 b = lds_f64x2(wbox + (uint32_t)(((c >> 1) * ((c >> 1) + 1) / 2 + (cp >> 1)) * 2048 + (c & 1) * 1024) +
                                                        wlrow + (wx0 ^ ((uint32_t)(cp & 1) << 6)));
                            dmma884(out[0][cq][0], out[0][cq][1], acc[0][cp][0], b.x);
                            dmma884(out[1][cq][0], out[1][cq][1], acc[1][cp][0], b.x);
                            dmma884(out[0][cq][0], out[0][cq][1], acc[0][cp][1], b.y);
                            dmma884(out[1][cq][0], out[1][cq][1], acc[1][cp][1], b.y);
                        }
                    }
                }
            }
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int cq = 0; cq < 4; ++cq) {
                    const int c = 4 * h + cq;
                    const int rl = 16 * wrow + 8 * f + g, cc = 8 * c + 2 * t;
                    if (rl < prow && cc < wk)
                        *reinterpret_cast<double2*>(Lb + (size_t)(r0 + rl) * ld + pc0 + cc) =
                            make_double2(-out[f][cq][0], -out[f][cq][1]);
                    sts_f64x2(Lh + (uint32_t)(c >> 1) * BOXB + (uint32_t)(2 * wl + f) * 1024 + lrow + (x0 ^ ((uint32_t)(c & 1) << 6)),
                              out[f][cq][0], out[f][cq][1]);
                }
        }
        // W_kk is dead: thread 0 may overwrite it for the next block
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.w_free);

        // ---- look-ahead: T_ii -= L_ik L_ik^T on this group's diagonal tile
        if (glook) {
            double acc2[2][8][2];
            const double* src = (k == 0 ? sigma : Lbuf) + bd.moff;
#pragma unroll
            for (int f = 0; f < 2; ++f)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int rl = 16 * wl + 8 * f + g, cc = 8 * c + 2 * t;
                    double2 v = make_double2(0.0, 0.0);
                    if (look && c <= 2 * wl + 1 && rl < tw && cc <= rl)
                        v = __ldcg(reinterpret_cast<const double2*>(src + (size_t)(trow0 + rl) * ld + trow0 + cc));
                    acc2[f][c][0] = v.x;
                    acc2[f][c][1] = v.y;
                }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
            if (look) {
#pragma unroll
                for (int f = 0; f < 2; ++f)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int rl = 16 * wl + 8 * f + g, cc = 8 * c + 2 * t;
                        double vx = acc2[f][c][0], vy = acc2[f][c][1];
                        if (k == 0 && trow0 + rl < bd.ms) {
                            if (cc == rl) vx += ridge;
                            if (cc + 1 == rl) vy += ridge;
                        }
                        acc2[f][c][0] = -vx;
                        acc2[f][c][1] = -vy;
                    }
                switch (wl) {
                    case 0: syrk_rows_pt<0>(Lh, lrow, x0, acc2); break;
                    case 1: syrk_rows_pt<1>(Lh, lrow, x0, acc2); break;
                    case 2: syrk_rows_pt<2>(Lh, lrow, x0, acc2); break;
                    default: syrk_rows_pt<3>(Lh, lrow, x0, acc2); break;
                }
#pragma unroll
                for (int f = 0; f < 2; ++f)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const int rl = 16 * wl + 8 * f + g, cc = 8 * c + 2 * t;
                        if (c <= 2 * wl + 1 && rl < tw && cc <= rl)
                            *reinterpret_cast<double2*>(Lb + (size_t)(trow0 + rl) * ld + trow0 + cc) =
                                make_double2(-acc2[f][c][0], -acc2[f][c][1]);
                    }
            }
        }
        // the ring boxes are free again (for the next item's chunks)
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.ring_free);

        if (fuse_end && item.y == 0 && r0 < bd.mp && wk == NB && j == n_my - 1) {
            __syncthreads();        // T_new is in L2 for the whole CTA; shared memory is free (last item: nothing in flight)
            diag_body<NT>(bd, item.x, k + 1, sigma, Lbuf, wbuf + wpar_next, ridge, status, reinterpret_cast<double*>(smem_al), []() {});
        }
    }
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
            {
                // x_k = W_kk^T v: the row of the diagonal tile is fetched with all (up to 16) loads in flight at once --
                // this matvec is on the per-panel dependent chain
                const double* row = Lb + (size_t)(pc0 + min(c, wk - 1)) * ld + pc0;
                double a[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int cp = c + q + 4 * i;
                    a[i] = (c < wk && cp < wk) ? row[cp] : 0.0;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int cp = c + q + 4 * i;
                    if (c < wk && cp < wk) p += ((cp == c) ? 1.0 / a[i] : a[i]) * v[cp];
                }
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
// Back substitution for big blocks (mp > 1024): a thread-block CLUSTER of 8 CTAs per block, right-looking and
// flag-driven, so that the dependent chain per panel is one 64x64 tile update + one 64x64 matvec with the stored
// W_kk^T and nothing else:
//   * CTA c owns the column panels j = c (mod 8) and keeps their running right-hand sides y_j in shared memory
//     (every update of y_j is done by its owner: no cross-CTA reduction);
//   * every CTA holds the whole solution vector x; the owner of panel k forms x_k = W_kk^T y_k, writes it into all
//     eight CTAs' shared memory through DSMEM and then raises flag k there (release at cluster scope);
//   * each CTA walks k downwards on its own: wait for flag k, fold row panel k into its panels j < k
//     (y_j -= L[k-rows][j-cols]^T x_k, 64 contiguous 512-byte row segments per tile, up to four tiles per pass);
//     the owner of panel k-1 folds that tile first and publishes x_k-1 before its other tiles.
// No cluster barrier inside the loop.  The column-panel form this replaces summed over ALL rows below per panel (two
// cluster barriers and a DSMEM reduction per panel, 14 us per panel); it was the critical path of the 8-GPU shards
// after the factorisation.
// ------------------------------------------------------------------------------------------
static constexpr int kBsCluster = 8;
static constexpr int kBsThreads = 512;

__device__ __forceinline__ void st_release_cluster_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.release.cluster.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_cluster_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.cluster.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
    return r;
}

__global__ void __cluster_dims__(kBsCluster, 1, 1) __launch_bounds__(kBsThreads)
backsolve_cluster_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ order,
                         const double* __restrict__ Lbuf, double inv_sqrt_n, double* __restrict__ beta_s,
                         double* __restrict__ beta_l, int32_t max_panels) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double smem[];
    const unsigned cr = cluster.block_rank();
    const BlockDesc bd = blocks[order[blockIdx.x / kBsCluster]];
    const int mp = bd.mp, ld = bd.ld;
    const double* Lb = Lbuf + bd.moff;
    const int max_owned = (max_panels + kBsCluster - 1) / kBsCluster;
    double* xall = smem;                           // [max_panels][64]  solution, every panel written by its owner (DSMEM)
    double* red = xall + max_panels * NB;          // [4][8][64]  per-row-group partial sums of up to four tiles
    double* yo = red + 32 * NB;                    // [max_owned][64] running right-hand sides of the owned panels
    uint32_t* flag = reinterpret_cast<uint32_t*>(yo + max_owned * NB);   // [max_panels] x_k has landed
    const int tid = threadIdx.x;
    const int K = (mp + NB - 1) / NB;
    // y = L^-1 z rides as matrix row mp
    for (int idx = tid; idx < max_owned * NB; idx += kBsThreads) {
        const int j = (idx >> 6) * kBsCluster + (int)cr, c = idx & 63;
        yo[idx] = (j < K && j * NB + c < mp) ? Lb[(size_t)mp * ld + j * NB + c] : 0.0;
    }
    for (int idx = tid; idx < max_panels; idx += kBsThreads) flag[idx] = 0u;
    __syncthreads();
    cluster.sync();                                // every CTA's flags are cleared before anyone publishes

    // x_k = W_kk^T y_k by the owner, delivered to every CTA, then flag k raised everywhere; beta written out
    auto solve_and_publish = [&](int k) {
        const int pc0 = k * NB, wk = min(NB, mp - pc0);
        const double* yk = yo + (k / kBsCluster) * NB;
        const int c = tid >> 3, q = tid & 7;       // 8 threads per entry
        double p = 0.0;
        {
            // all (up to eight) loads of a thread are issued together: this matvec sits on the dependent chain
            const double* row = Lb + (size_t)(pc0 + min(c, wk - 1)) * ld + pc0;   // diagonal at row[c], W^T[c][cp] at row[cp], cp > c
            double a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int cp = c + q + 8 * i;
                a[i] = (c < wk && cp < wk) ? row[cp] : 0.0;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int cp = c + q + 8 * i;
                if (c < wk && cp < wk) p += ((cp == c) ? 1.0 / a[i] : a[i]) * yk[cp];
            }
        }
        p += __shfl_xor_sync(0xffffffffu, p, 1);
        p += __shfl_xor_sync(0xffffffffu, p, 2);
        p += __shfl_xor_sync(0xffffffffu, p, 4);
        cluster.map_shared_rank(xall, q)[k * NB + c] = (c < wk) ? p : 0.0;    // thread (c, q) delivers x_k[c] to CTA q
        if (q == 0 && c < wk) {
            const int j = pc0 + c;
            if (j < bd.m) {
                const double b = p * inv_sqrt_n;
                if (j < bd.ms) beta_s[bd.out_s + j] = b;
                else beta_l[bd.out_l + (j - bd.ms)] = b;
            }
        }
        asm volatile("fence.acq_rel.cluster;" ::: "memory");
        __syncthreads();                           // all 512 deliveries are ordered before the flags
        if (tid < kBsCluster) st_release_cluster_u32(mapa_u32(smem_u32(flag + k), (uint32_t)tid), 1u);
    };
    // y_j -= L[panel k rows][panel j cols]^T x_k for up to four owned panels j0, j0-8, ... in ONE pass
    auto fold_batch = [&](int k, int j0, int ntile) {
        const int pc0 = k * NB, wk = min(NB, mp - pc0);
        const double* x = xall + k * NB;
        const int rg = tid >> 6, c = tid & 63;     // 8 row groups x 64 columns
        double p[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j0 - u * kBsCluster;
            if (u < ntile && j >= 0) {
                const double* col = Lb + (size_t)pc0 * ld + j * NB + c;
                double a[8];                                   // eight independent loads in flight per tile
#pragma unroll
                for (int i = 0; i < 8; ++i) { const int r = rg + 8 * i; a[i] = (r < wk) ? col[(size_t)r * ld] : 0.0; }
#pragma unroll
                for (int i = 0; i < 8; ++i) { const int r = rg + 8 * i; if (r < wk) p[u] += a[i] * x[r]; }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) red[(u * 8 + rg) * NB + c] = p[u];
        __syncthreads();
        if (tid < 4 * NB) {
            const int u = tid >> 6, c2 = tid & 63, j = j0 - u * kBsCluster;
            if (u < ntile && j >= 0) {
                double sacc = 0.0;
#pragma unroll
                for (int g2 = 0; g2 < 8; ++g2) sacc += red[(u * 8 + g2) * NB + c2];  // fixed order
                yo[(j / kBsCluster) * NB + c2] -= sacc;
            }
        }
        __syncthreads();
    };

    if ((K - 1) % kBsCluster == (int)cr) solve_and_publish(K - 1);
    for (int k = K - 1; k >= 1; --k) {
        if (tid == 0) { while (ld_acquire_cluster_u32(flag + k) == 0u) { } }
        __syncthreads();                           // x_k is in xall
        // owned panels below k, highest first: j = k-1 (if mine) is on the critical path and goes alone
        int j = k - 1 - (((k - 1) - (int)cr) % kBsCluster + kBsCluster) % kBsCluster;   // largest j <= k-1 with j = cr (mod 8)
        if (j == k - 1) {
            fold_batch(k, j, 1);
            solve_and_publish(k - 1);
            j -= kBsCluster;
        }
        for (; j >= 0; j -= 4 * kBsCluster) fold_batch(k, j, 4);
    }
    cluster.sync();                                // nobody exits while its shared memory may still be written
}

cudaError_t chol_configure() {
    cudaError_t e = cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_DIAG);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(chol_panel_tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TpCfg<2>::SMEM);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(chol_panel_tma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TpCfg<1>::SMEM);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_CHOL);
}

// panel CTAs that fit on one SM, by macro-tile height (the plan sizes split-K and the deferral rule with it)
int chol_panel_ctas_per_sm(int32_t tile_rows) { return tile_rows == 64 ? TpCfg<1>::MINB : 2; }

// TMA panel step (see chol_panel_tma_kernel).  tile_rows = 64 or 128 (rows per item, as the plan built the list);
// n_single leading items get a CTA each, the rest `tpc` consecutive items per CTA; pf = L2 prefetch distance in chunks.
cudaError_t launch_chol_panel_tma(int32_t tile_rows, const BlockDesc* blocks, const int4* items, int32_t n_items, int32_t n_single,
                                  int32_t tpc, const int32_t* diag_items, int32_t n_diag_first, int32_t k, const CUtensorMap* lmaps,
                                  int32_t perm, int32_t pf, const double* sigma, double* L, double* wbuf, int64_t wstride,
                                  bool fuse_end, double ridge, double* scratch, int32_t* counters, int32_t group_base,
                                  int32_t* status, int32_t* dflag, bool pdl, cudaStream_t st) {
    if (n_items + n_diag_first == 0) return cudaSuccess;
    const int64_t wpar = (k & 1) * wstride, wnext = ((k + 1) & 1) * wstride;
    n_single = std::min(n_single, n_items);
    tpc = std::max(tpc, 1);
    const int rest = n_items - n_single;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_diag_first + n_single + (rest + tpc - 1) / tpc));
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (tile_rows == 64) {
        cfg.blockDim = dim3(TpCfg<1>::NT);
        cfg.dynamicSmemBytes = TpCfg<1>::SMEM;
        return cudaLaunchKernelEx(&cfg, chol_panel_tma_kernel<1>, blocks, items, n_items, n_single, tpc, diag_items, n_diag_first, k,
                                  lmaps, perm, pf, sigma, L, wbuf, wpar, wnext, ridge, scratch, counters, group_base, status, dflag,
                                  (int32_t)(fuse_end ? 1 : 0));
    }
    cfg.blockDim = dim3(TpCfg<2>::NT);
    cfg.dynamicSmemBytes = TpCfg<2>::SMEM;
    return cudaLaunchKernelEx(&cfg, chol_panel_tma_kernel<2>, blocks, items, n_items, n_single, tpc, diag_items, n_diag_first, k,
                              lmaps, perm, pf, sigma, L, wbuf, wpar, wnext, ridge, scratch, counters, group_base, status, dflag,
                              (int32_t)(fuse_end ? 1 : 0));
}

// W_k (the inverse of the diagonal tile of panel k) lives in wbuf[(k & 1) * wstride + block * 64 * 64]: two parities,
// because with the fused diagonal the step-k launch reads W_k while its macro-tile-0 CTAs already write W_k+1.
cudaError_t launch_chol_diag(const BlockDesc* blocks, const int32_t* items, int32_t n_items, int32_t k,
                             const double* sigma, double* L, double* wbuf, int64_t wstride, double ridge,
                             int32_t* status, cudaStream_t st) {
    if (n_items == 0) return cudaSuccess;
    chol_diag_kernel<<<n_items, CHOL_THREADS, SMEM_DIAG, st>>>(blocks, items, k, sigma, L, wbuf + (k & 1) * wstride, ridge,
                                                              status);
    return cudaGetLastError();
}
// diag_first: the n_diag blocks of `diag_items` get their diagonal tile of panel k factored by extra CTAs of this
// launch (step k-1 deferred it); fuse_end: macro tile 0 factors the diagonal tile of panel k+1 at its end.
cudaError_t launch_chol_panel(const BlockDesc* blocks, const int4* items, int32_t n_items, const int32_t* diag_items,
                              int32_t n_diag_first, int32_t k, const double* sigma, double* L, double* wbuf, int64_t wstride,
                              bool fuse_end, double ridge, double* scratch, int32_t* counters, int32_t group_base,
                              int32_t* status, int32_t* dflag, bool pdl, cudaStream_t st) {
    if (n_items + n_diag_first == 0) return cudaSuccess;
    const int64_t wpar = (k & 1) * wstride, wnext = ((k + 1) & 1) * wstride;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(n_items + n_diag_first));
    cfg.blockDim = dim3(CHOL_THREADS);
    cfg.dynamicSmemBytes = SMEM_CHOL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, chol_panel_kernel, blocks, items, diag_items, n_diag_first, k, sigma, L, wbuf, wpar, wnext, ridge,
                              scratch, counters, group_base, status, dflag, (int32_t)(fuse_end ? 1 : 0));
}
// Back substitution of `n_blocks` blocks listed in `order`: an 8-CTA cluster per block for the big size
// classes (mp > 1024), one 256-thread CTA per block otherwise.
cudaError_t launch_backsolve(const BlockDesc* blocks, const int32_t* order, int32_t n_blocks, bool big,
                             const double* L, double inv_sqrt_n, double* beta_s, double* beta_l, int32_t max_mp,
                             cudaStream_t st) {
    if (n_blocks == 0) return cudaSuccess;
    if (big) {
        const int K = (max_mp + NB - 1) / NB;
        const int max_owned = (K + kBsCluster - 1) / kBsCluster;
        const size_t smem = (size_t)(K * NB + 32 * NB + max_owned * NB) * sizeof(double) + (size_t)K * sizeof(uint32_t) + 16;
        cudaError_t e = cudaFuncSetAttribute(backsolve_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        backsolve_cluster_kernel<<<n_blocks * kBsCluster, kBsThreads, smem, st>>>(blocks, order, L, inv_sqrt_n, beta_s, beta_l, K);
    } else {
        const size_t smem = (size_t)(1024 + 9 * NB) * sizeof(double);
        backsolve_kernel<256><<<n_blocks, 256, smem, st>>>(blocks, order, L, inv_sqrt_n, beta_s, beta_l);
    }
    return cudaGetLastError();
}

}  // namespace dbslmm
