// pcg.cu -- reference-faithful block solver: Jacobi-preconditioned conjugate gradients with the
// reference's own stopping rule, and estBlock's algebra verbatim.
//
// Follows DBSLMMFIT::PCGv / PCGm (reference scr/dbslmmfit.cpp:629-678) and both estBlock overloads
// (:680-738 large+small, :740-770 small only): Minv = 1/diag(A) with a 1e-4 guard for zero diagonals,
// x0 = 0, loop while ||r||_2 > 1e-7 and iter < 1000, then the Schur complement, the large-effect solve
// and the reference's cancellation form for beta_s.  It exists so that the GPU path can reproduce the
// reference's TRUNCATED answers (its PCG stops at an absolute residual of 1e-7); the default solver is
// the exact Cholesky in chol.cu.  Two PCG implementations that differ only in summation order agree to
// ~1e-10 while iteration counts coincide and to the truncation error (few 1e-8) when a count flips at
// the threshold (tests/test_oracle.py), so this mode is about fidelity, not speed: one 1024-thread CTA
// per block, right-hand sides one after the other like PCGm, every reduction in a fixed order.
#include "common.cuh"
#include "kernels.h"

namespace dbslmm {

static constexpr int kPcgThreads = 1024;
static constexpr int kPcgWarps = kPcgThreads / 32;

// deterministic block-wide sum (fixed tree), result broadcast to all threads
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = (lane < kPcgWarps) ? red[lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return s;
}

// y = (A + ridge*I) p for the n x n matrix at A (row stride ld); p staged in shared memory
__device__ __forceinline__ void matvec(const double* __restrict__ A, int ld, int n, double ridge, const double* ps,
                                       double* __restrict__ y) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = warp; i < n; i += kPcgWarps) {
        const double* row = A + (size_t)i * ld;
        double s = 0.0;
        for (int j = lane; j < n; j += 32) s += row[j] * ps[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) y[i] = s + ridge * ps[i];
    }
}

// DBSLMMFIT::PCGv (dbslmmfit.cpp:629-668).  b may alias nothing in the work vectors.  Returns iterations.
__device__ int pcgv(const double* __restrict__ A, int ld, int n, double ridge, const double* __restrict__ b,
                    double* __restrict__ x, double* __restrict__ r, double* __restrict__ z, double* __restrict__ p,
                    double* __restrict__ Ap, double* ps, double* red) {
    const int tid = threadIdx.x;
    if (n == 0) return 0;
    double loc = 0.0;
    for (int i = tid; i < n; i += kPcgThreads) {
        double d = A[(size_t)i * ld + i] + ridge;
        if (d == 0) d = 1e-4;                              // :632-635
        const double minv = 1.0 / d;                       // :636
        const double bi = b[i];
        x[i] = 0.0;                                        // :638
        r[i] = bi;                                         // :642
        z[i] = minv * bi;                                  // :643
        p[i] = z[i];                                       // :644
        Ap[i] = 0.0;
        loc += bi * bi;
    }
    double sumr2 = sqrt(block_sum(loc, red));              // :646
    int iter = 0;
    while (sumr2 > 1e-7 && iter < 1000) {                  // :648
        iter += 1;
        __syncthreads();
        for (int i = tid; i < n; i += kPcgThreads) ps[i] = p[i];
        __syncthreads();
        matvec(A, ld, n, ridge, ps, Ap);                   // :651
        __syncthreads();
        double rz = 0.0, pAp = 0.0;
        for (int i = tid; i < n; i += kPcgThreads) { rz += r[i] * z[i]; pAp += ps[i] * Ap[i]; }
        rz = block_sum(rz, red);
        pAp = block_sum(pAp, red);
        const double a = rz / pAp;                         // :653
        double z1r1 = 0.0, rr = 0.0;
        for (int i = tid; i < n; i += kPcgThreads) {
            double d = A[(size_t)i * ld + i] + ridge;
            if (d == 0) d = 1e-4;
            const double minv = 1.0 / d;
            x[i] = x[i] + a * ps[i];                       // :655
            const double r1 = r[i] - a * Ap[i];            // :656
            const double z1 = minv * r1;                   // :657
            r[i] = r1;
            z[i] = z1;
            z1r1 += z1 * r1;
            rr += r1 * r1;
        }
        z1r1 = block_sum(z1r1, red);
        rr = block_sum(rr, red);
        const double bet = z1r1 / rz;                      // :658 (z.r of the previous iterate)
        for (int i = tid; i < n; i += kPcgThreads) p[i] = z[i] + bet * ps[i];   // :659
        sumr2 = sqrt(rr);                                  // :662
    }
    __syncthreads();
    return iter;
}

__global__ void __launch_bounds__(kPcgThreads, 1)
pcg_kernel(const PcgArgs a) {
    extern __shared__ __align__(16) double ps[];           // [max ms] search direction
    __shared__ double red[kPcgWarps];
    const int blk = a.order[blockIdx.x];
    const BlockDesc bd = a.blocks[blk];
    if (bd.m == 0) return;
    const int tid = threadIdx.x;
    const int ms = bd.ms, ml = bd.m - bd.ms, ld = bd.ld;
    const double* S = a.sigma + bd.moff;                   // full symmetric Sigma (small rows first)
    const double* zrow = S + (size_t)bd.mp * ld;           // z_s then z_l
    double* w = a.work + a.work_off[blk];
    double *x = w, *r = x + ms, *z = r + ms, *p = z + ms, *Ap = p + ms, *u = Ap + ms, *tvec = u + ms;
    double* W = tvec + ms;                                 // [ml][ms]  column c of A^-1 Sigma_sl
    double* Sc = W + (size_t)ms * ml;                      // [ml][ml]  Schur complement (row-major)
    double* rhs = Sc + (size_t)ml * ml;                    // [ml]
    double* bl = rhs + ml;                                 // [ml]
    double* xs = bl + ml;                                  // [ml] x 5 work vectors of the small system
    const double sq = sqrt(a.n_obs), dn = a.n_obs;
    int itmax = 0, sing = 0;
    auto track = [&](int it) { itmax = max(itmax, it); if (it >= 1000) sing = 1; };

    if (ml == 0) {
        // ---- small-only overload (:740-770)
        track(pcgv(S, ld, ms, a.ridge, zrow, u, r, z, p, Ap, ps, red));                 // :760
        for (int i = tid; i < ms; i += kPcgThreads) ps[i] = u[i];
        __syncthreads();
        matvec(S, ld, ms, 0.0, ps, tvec);                                               // :762 (ridge removed :761)
        __syncthreads();
        for (int i = tid; i < ms; i += kPcgThreads) a.beta_s[bd.out_s + i] = sq * a.sigma_s * (zrow[i] - tvec[i]);   // :763-764
    } else {
        // ---- large + small overload (:680-738)
        for (int c = 0; c < ml; ++c) {                                                   // PCGm :670-678, call :713
            // right-hand side = column c of Sigma_sl = row (ms + c) of Sigma restricted to the small columns
            track(pcgv(S, ld, ms, a.ridge, S + (size_t)(ms + c) * ld, W + (size_t)c * ms, r, z, p, Ap, ps, red));
        }
        // S = Sigma_ll - Sigma_ls W   (:714-715), entry (a2, b2)
        for (int e = tid >> 5; e < ml * ml; e += kPcgWarps) {
            const int a2 = e / ml, b2 = e - a2 * ml;
            const double* sls = S + (size_t)(ms + a2) * ld;
            const double* wc = W + (size_t)b2 * ms;
            double s = 0.0;
            for (int i = tid & 31; i < ms; i += 32) s += sls[i] * wc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((tid & 31) == 0) Sc[(size_t)a2 * ml + b2] = -s + S[(size_t)(ms + a2) * ld + ms + b2];
        }
        track(pcgv(S, ld, ms, a.ridge, zrow, u, r, z, p, Ap, ps, red));                  // :716
        for (int e = tid >> 5; e < ml; e += kPcgWarps) {                                  // :717-718
            const double* sls = S + (size_t)(ms + e) * ld;
            double s = 0.0;
            for (int i = tid & 31; i < ms; i += 32) s += sls[i] * u[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((tid & 31) == 0) rhs[e] = -s + zrow[ms + e];
        }
        __syncthreads();
        track(pcgv(Sc, ml, ml, 0.0, rhs, bl, xs, xs + ml, xs + 2 * ml, xs + 3 * ml, ps, red));   // :719
        for (int e = tid; e < ml; e += kPcgThreads) bl[e] = bl[e] / sq;                   // :720
        __syncthreads();
        // t = sqrt(N) u - N W beta_l   (:723-725)
        for (int i = tid; i < ms; i += kPcgThreads) {
            double s = 0.0;
            for (int c = 0; c < ml; ++c) s += W[(size_t)c * ms + i] * bl[c];
            ps[i] = sq * u[i] - dn * s;
        }
        __syncthreads();
        matvec(S, ld, ms, 0.0, ps, tvec);                                                 // :727 (ridge removed :726)
        __syncthreads();
        for (int i = tid; i < ms; i += kPcgThreads) {                                     // :728-729
            double s = 0.0;
            for (int c = 0; c < ml; ++c) s += S[(size_t)(ms + c) * ld + i] * bl[c];
            a.beta_s[bd.out_s + i] = a.sigma_s * (sq * zrow[i] - dn * s - tvec[i]);
        }
        for (int e = tid; e < ml; e += kPcgThreads) a.beta_l[bd.out_l + e] = bl[e];
    }
    if (tid == 0) {
        a.iters[blk] = itmax;
        if (sing) atomicOr(&a.status[blk], 2);             // "ERROR: Matrix is Singular!" (:664-666)
    }
}

cudaError_t launch_pcg(const PcgArgs& a, int32_t max_ms, cudaStream_t st) {
    if (a.n_blocks == 0) return cudaSuccess;
    const size_t smem = sizeof(double) * (size_t)(max_ms + 8);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(pcg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    pcg_kernel<<<a.n_blocks, kPcgThreads, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace dbslmm
