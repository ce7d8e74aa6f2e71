// decode.cu -- genotype decoder (K1) and per-SNP statistics (the MAF pre-pass, a3).
//
// Replaces IO::readSNPIm + SNPPROC::nomalizeVec (reference scr/dtpr.cpp:285-380): instead
// of one ifstream::read per byte and an FP64 column per SNP, each warp stages one .bed row
// into shared memory with a 1-D TMA bulk copy (cp.async.bulk, double buffered on
// mbarriers), unpacks sixteen 2-bit codes per lane with bit arithmetic and writes int8
// allele counts with coalesced 16-byte stores.  Standardisation is NOT applied here: it is
// folded into the Gram epilogue as exact integer arithmetic (gram.cu); the decoder only
// emits the per-row scale factor.
//
// Code map (dtpr.cpp:329-350): 2-bit v = b0 + 2*b1, low bits first within a byte:
//   v=0 -> 2, v=2 -> 1, v=3 -> 0, v=1 -> missing (genotype plane 0, mask plane 0).
#include <algorithm>
#include "common.cuh"
#include "kernels.h"

namespace dbslmm {

static constexpr int kWarpsPerCta = 8;
// staged rows per warp in the decoder: 4 (three row copies in flight per warp), 2 for very long rows (n_ref > ~25,000)

struct RowStage {
    // returns byte offset of the row inside the staged buffer
    __device__ static __forceinline__ uint32_t issue(const uint8_t* bed, int64_t row, int32_t pitch,
                                                     uint8_t* buf, uint64_t* bar, int lane) {
        const int64_t s = row * (int64_t)pitch;
        const int64_t a0 = s & ~(int64_t)15;
        const uint32_t nbytes = (uint32_t)(((s + pitch + 15) & ~(int64_t)15) - a0);
        if (lane == 0) {
            mbar_expect_tx(bar, nbytes);
            bulk_g2s(buf, bed + a0, nbytes, bar);
        }
        return (uint32_t)(s - a0);
    }
};

// 32-bit word `i` of a row that starts `off` bytes into the 16-byte aligned staging buffer
__device__ __forceinline__ uint32_t row_word(const uint32_t* w32, uint32_t off, int i) {
    const uint32_t idx = (off >> 2) + i;
    const uint32_t lo = w32[idx], hi = w32[idx + 1];
    return __funnelshift_r(lo, hi, (off & 3u) * 8u);
}

// mask with the low 2*k bits set, k = number of valid samples in word i (0..16)
__device__ __forceinline__ uint32_t valid_bits(int n_ref, int i) {
    int k = n_ref - 16 * i;
    k = k < 0 ? 0 : (k > 16 ? 16 : k);
    return k == 16 ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
}

// ------------------------------------------------------------------------------------------
// Per-SNP statistics over the whole .bed: non-missing count, allele sum, sum of squares.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerCta * 32)
snp_stats_kernel(const uint8_t* __restrict__ bed, int64_t n_snp, int32_t n_ref, int32_t pitch,
                 int32_t buf_bytes, SnpStat* __restrict__ stats) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[kWarpsPerCta][2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* buf0 = smem + (size_t)warp * 2 * buf_bytes;
    uint8_t* bufs[2] = {buf0, buf0 + buf_bytes};
    if (lane == 0) { mbar_init(&bars[warp][0], 1); mbar_init(&bars[warp][1], 1); }
    mbar_fence_init();
    __syncwarp();

    const int wpc = blockDim.x >> 5;          // warps per CTA (8 unless a row is too long for 16 staging buffers)
    const int64_t gw = (int64_t)blockIdx.x * wpc + warp;
    const int64_t stride = (int64_t)gridDim.x * wpc;
    const int nwords = (pitch + 3) >> 2;
    uint32_t phase[2] = {0, 0};
    uint32_t off[2] = {0, 0};
    int cur = 0;
    int64_t row = gw;
    if (row < n_snp) off[0] = RowStage::issue(bed, row, pitch, bufs[0], &bars[warp][0], lane);
    for (; row < n_snp; row += stride) {
        const int64_t nxt = row + stride;
        if (nxt < n_snp) off[cur ^ 1] = RowStage::issue(bed, nxt, pitch, bufs[cur ^ 1], &bars[warp][cur ^ 1], lane);
        mbar_wait(&bars[warp][cur], phase[cur]);
        phase[cur] ^= 1;
        const uint32_t* w32 = reinterpret_cast<const uint32_t*>(bufs[cur]);
        int c0 = 0, c1 = 0, c2 = 0;
        for (int i = lane; i < nwords; i += 32) {
            uint32_t w = row_word(w32, off[cur], i);
            w |= ~valid_bits(n_ref, i);                // samples past n_ref -> code 3 (counts nothing)
            const uint32_t lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
            c0 += __popc(~(lo | hi) & 0x55555555u);    // code 0 -> allele count 2
            c1 += __popc(lo & ~hi);                    // code 1 -> missing
            c2 += __popc(hi & ~lo);                    // code 2 -> allele count 1
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, o);
            c1 += __shfl_xor_sync(0xffffffffu, c1, o);
            c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        }
        if (lane == 0) {
            SnpStat s;
            s.n_nonmiss = n_ref - c1;
            s.sum = 2 * c0 + c2;
            s.sumsq = 4 * c0 + c2;
            s.pad = 0;
            stats[row] = s;
        }
        __syncwarp();
        cur ^= 1;
    }
}

// sixteen 2-bit codes of word w -> sixteen allele-count bytes (MASK = false: 0->2, 2->1, 1/3->0) or call-mask bytes
// (MASK = true: 0 only for code 1 = missing).  Each 2-bit code is spread into its own nibble (three shift+LOP3
// steps per 8 codes) and used as a PRMT byte selector into a 4-entry table held in one register.
template <bool MASK>
__device__ __forceinline__ uint2 expand8(uint32_t h) {                 // h: 8 codes in the low 16 bits
    constexpr uint32_t lut = MASK ? 0x01010001u : 0x00010002u;         // byte c = value of code c
    uint32_t t = (h | (h << 8)) & 0x00FF00FFu;
    t = (t | (t << 4)) & 0x0F0F0F0Fu;
    t = (t | (t << 2)) & 0x33333333u;
    return make_uint2(__byte_perm(lut, 0u, t), __byte_perm(lut, 0u, t >> 16));
}
template <bool MASK>
__device__ __forceinline__ uint4 expand16(uint32_t w) {
    const uint2 lo = expand8<MASK>(w & 0xFFFFu), hi = expand8<MASK>(w >> 16);
    return make_uint4(lo.x, lo.y, hi.x, hi.y);
}

// ------------------------------------------------------------------------------------------
// Decoder: one warp per SNP row of the plan.  The warp stages the .bed row ONCE, counts its genotypes (the per-SNP
// statistics of the standardisation: no separate pass), writes the int8 allele-count row, and -- only if the SNP has
// missing calls, or the mask row at that position was written for an earlier fit -- the int8 call-mask row.
//   row_src[g]  .bed row of SNP row g          row_crow[g]  its genotype code row      row_mrow[g]  its mask code row
//   dirty[c]    code row c holds a non-default mask (device state that lives across fits; starts all-ones)
// Invariant: a mask row whose dirty byte is 0 holds the default pattern (1 for every real sample, 0 for padding), so a
// panel without missing calls never pays for the mask plane, and the correlation builder can pick the one- or the
// four-plane path per block from a flag computed on the device (block_flags_kernel).
// ------------------------------------------------------------------------------------------
template <int kDecRing>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
decode_rows_kernel(const uint8_t* __restrict__ bed, int32_t n_ref, int32_t pitch, int32_t n_pad,
                   int32_t buf_bytes, const uint32_t* __restrict__ row_src, const int32_t* __restrict__ row_crow,
                   const int32_t* __restrict__ row_mrow, int64_t g0, int64_t n_rows, double tau,
                   int8_t* __restrict__ codes, uint8_t* __restrict__ dirty, int32_t* __restrict__ rowN,
                   int32_t* __restrict__ rowS, double* __restrict__ rowR, double2* __restrict__ rowC, const int32_t* __restrict__ gate) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[kWarpsPerCta][kDecRing];
    // gate (fused-Gram pipeline): int8 rows are needed only by blocks with missing calls; *gate == 0 says there are none
    if (gate != nullptr && *gate == 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* buf0 = smem + (size_t)warp * kDecRing * buf_bytes;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kDecRing; ++i) mbar_init(&bars[warp][i], 1);
    }
    mbar_fence_init();
    __syncwarp();

    const int wpc = blockDim.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * wpc + warp;
    const int64_t stride = (int64_t)gridDim.x * wpc;
    const int nwords = (pitch + 3) >> 2;     // input words holding real samples
    const int nout = n_pad >> 4;             // 16-byte output chunks per row
    // ring of kDecRing staged rows per warp: kDecRing - 1 TMA row copies in flight while one row is expanded.  The
    // source index of the row staged NEXT iteration is loaded one iteration ahead, so no global load sits on the
    // loop's critical path.
    uint32_t phase = 0;                      // bit i = parity of buffer i
    uint32_t off[kDecRing];
    int64_t row = gw;                        // relative to g0
#pragma unroll
    for (int i = 0; i < kDecRing; ++i) off[i] = 0;
#pragma unroll
    for (int i = 0; i < kDecRing - 1; ++i) {
        const int64_t r = row + (int64_t)i * stride;
        if (r < n_rows) off[i] = RowStage::issue(bed, (int64_t)row_src[g0 + r], pitch, buf0 + i * buf_bytes, &bars[warp][i], lane);
    }
    uint32_t pre = 0;                        // row_src of the row that will be staged in the current iteration
    {
        const int64_t r = row + (int64_t)(kDecRing - 1) * stride;
        if (r < n_rows) pre = row_src[g0 + r];
    }
    int cur = 0;
    for (; row < n_rows; row += stride) {
        {
            const int64_t nxt = row + (int64_t)(kDecRing - 1) * stride;
            const int nb = (cur + kDecRing - 1) % kDecRing;
            const uint32_t issue_src = pre;
            const int64_t nxt2 = nxt + stride;
            if (nxt2 < n_rows) pre = row_src[g0 + nxt2];     // consumed next iteration
            if (nxt < n_rows) {
                const uint32_t o = RowStage::issue(bed, (int64_t)issue_src, pitch, buf0 + nb * buf_bytes, &bars[warp][nb], lane);
#pragma unroll
                for (int i = 0; i < kDecRing; ++i) if (i == nb) off[i] = o;
            }
        }
        uint32_t off_cur = 0;
#pragma unroll
        for (int i = 0; i < kDecRing; ++i) if (i == cur) off_cur = off[i];
        const int64_t g = g0 + row;
        const int32_t crow = row_crow[g], mrow = row_mrow[g];
        const uint8_t was_dirty = dirty[mrow];
        mbar_wait(&bars[warp][cur], (phase >> cur) & 1u);
        phase ^= 1u << cur;
        const uint32_t* w32 = reinterpret_cast<const uint32_t*>(buf0 + cur * buf_bytes);
        uint4* out = reinterpret_cast<uint4*>(codes + (size_t)crow * n_pad);
        const int nfull = n_ref >> 4;             // words whose 16 samples are all real
        int c0 = 0, c1 = 0, c2 = 0;
        for (int i = lane; i < nout; i += 32) {
            uint4 o = make_uint4(0u, 0u, 0u, 0u);
            if (i < nwords) {
                uint32_t w = row_word(w32, off_cur, i);
                uint32_t ws = w;                          // for the counts: samples past n_ref -> code 3 (counts nothing)
                if (i >= nfull) {
                    const uint32_t vb = valid_bits(n_ref, i);
                    ws = w | ~vb;
                    w = (w & vb) | (0x55555555u & ~vb);   // samples past n_ref -> "missing": 0 in both planes
                }
                o = expand16<false>(w);
                const uint32_t lo = ws & 0x55555555u, hi = (ws >> 1) & 0x55555555u;
                c0 += __popc(~(lo | hi) & 0x55555555u);    // code 0 -> allele count 2
                c1 += __popc(lo & ~hi);                    // code 1 -> missing
                c2 += __popc(hi & ~lo);                    // code 2 -> allele count 1
            }
            out[i] = o;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, o);
            c1 += __shfl_xor_sync(0xffffffffu, c1, o);
            c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        }
        if (c1 > 0 || was_dirty) {
            // the call-mask plane of this SNP (second pass over the staged row; rare on real reference panels)
            uint4* outm = reinterpret_cast<uint4*>(codes + (size_t)mrow * n_pad);
            for (int i = lane; i < nout; i += 32) {
                uint4 o = make_uint4(0u, 0u, 0u, 0u);
                if (i < nwords) {
                    uint32_t w = row_word(w32, off_cur, i);
                    if (i >= nfull) { const uint32_t vb = valid_bits(n_ref, i); w = (w & vb) | (0x55555555u & ~vb); }
                    o = expand16<true>(w);
                }
                outm[i] = o;
            }
        }
        if (lane == 0) {
            if ((c1 > 0) != (was_dirty != 0)) dirty[mrow] = (c1 > 0) ? 1 : 0;
            const int32_t nn = n_ref - c1, sum = 2 * c0 + c2, sumsq = 4 * c0 + c2;
            const double ni = (double)nn;
            // d = n_i * sum g^2 - (sum g)^2  (exact integer), r = sqrt(tau (n-1) / (n n_i d))
            const double d = ni * (double)sumsq - (double)sum * (double)sum;
            const double n = (double)n_ref;
            rowN[g] = nn;
            rowS[g] = sum;
            const double rr = sqrt(tau * (n - 1.0) / (n * ni * d));
            rowR[g] = rr;
            rowC[g] = make_double2((double)sum, rr);      // {S_i, r_i} as one 16-byte record: the Gram epilogues stage it with cp.async
        }
        __syncwarp();
        cur = (cur + 1 == kDecRing) ? 0 : cur + 1;
    }
}

// ------------------------------------------------------------------------------------------
// Packer: the decoder's front half for the fused correlation builder (gram.cu: gram_packed_kernel).  One warp per SNP row
// of the plan stages the .bed row (same TMA bulk ring), counts its genotypes (per-SNP statistics) and writes the row as
// 2-BIT codes again -- but gathered into plan order, 16-byte aligned (pitch n_pad / 4) and with the samples past n_ref set
// to code 3 (zero copies) -- so the correlation builder can fetch [128 rows x 128 samples] operand tiles with ONE 2-D TMA
// load each and unpack them in shared memory.  4 x less operand traffic through L2 than int8 rows; a quarter of the bytes
// written here.
// ------------------------------------------------------------------------------------------
template <int kDecRing>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
pack_rows_kernel(const uint8_t* __restrict__ bed, int32_t n_ref, int32_t pitch, int32_t n_pad, int32_t buf_bytes,
                 const uint32_t* __restrict__ row_src, int64_t g0, int64_t n_rows, double tau,
                 uint32_t* __restrict__ packed, int32_t* __restrict__ rowN, int32_t* __restrict__ rowS,
                 double* __restrict__ rowR, double2* __restrict__ rowC) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[kWarpsPerCta][kDecRing];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* buf0 = smem + (size_t)warp * kDecRing * buf_bytes;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kDecRing; ++i) mbar_init(&bars[warp][i], 1);
    }
    mbar_fence_init();
    __syncwarp();
    const int wpc = blockDim.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * wpc + warp;
    const int64_t stride = (int64_t)gridDim.x * wpc;
    const int nwords = (pitch + 3) >> 2;     // input words holding real samples
    const int nout = n_pad >> 4;             // 32-bit output words per row (16 samples each)
    const int nfull = n_ref >> 4;
    uint32_t phase = 0;
    uint32_t off[kDecRing];
    int64_t row = gw;
#pragma unroll
    for (int i = 0; i < kDecRing; ++i) off[i] = 0;
#pragma unroll
    for (int i = 0; i < kDecRing - 1; ++i) {
        const int64_t r = row + (int64_t)i * stride;
        if (r < n_rows) off[i] = RowStage::issue(bed, (int64_t)row_src[g0 + r], pitch, buf0 + i * buf_bytes, &bars[warp][i], lane);
    }
    uint32_t pre = 0;
    {
        const int64_t r = row + (int64_t)(kDecRing - 1) * stride;
        if (r < n_rows) pre = row_src[g0 + r];
    }
    int cur = 0;
    for (; row < n_rows; row += stride) {
        {
            const int64_t nxt = row + (int64_t)(kDecRing - 1) * stride;
            const int nb = (cur + kDecRing - 1) % kDecRing;
            const uint32_t issue_src = pre;
            const int64_t nxt2 = nxt + stride;
            if (nxt2 < n_rows) pre = row_src[g0 + nxt2];
            if (nxt < n_rows) {
                const uint32_t o = RowStage::issue(bed, (int64_t)issue_src, pitch, buf0 + nb * buf_bytes, &bars[warp][nb], lane);
#pragma unroll
                for (int i = 0; i < kDecRing; ++i) if (i == nb) off[i] = o;
            }
        }
        uint32_t off_cur = 0;
#pragma unroll
        for (int i = 0; i < kDecRing; ++i) if (i == cur) off_cur = off[i];
        const int64_t g = g0 + row;
        mbar_wait(&bars[warp][cur], (phase >> cur) & 1u);
        phase ^= 1u << cur;
        const uint32_t* w32 = reinterpret_cast<const uint32_t*>(buf0 + cur * buf_bytes);
        uint32_t* out = packed + (size_t)g * nout;
        int c0 = 0, c1 = 0, c2 = 0;
        for (int i = lane; i < nout; i += 32) {
            uint32_t w = 0xFFFFFFFFu;                     // 16 x code 3: samples past the panel count nothing, unpack to 0
            if (i < nwords) {
                w = row_word(w32, off_cur, i);
                if (i >= nfull) w |= ~valid_bits(n_ref, i);
                const uint32_t lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
                c0 += __popc(~(lo | hi) & 0x55555555u);    // code 0 -> allele count 2
                c1 += __popc(lo & ~hi);                    // code 1 -> missing
                c2 += __popc(hi & ~lo);                    // code 2 -> allele count 1
            }
            out[i] = w;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            c0 += __shfl_xor_sync(0xffffffffu, c0, o);
            c1 += __shfl_xor_sync(0xffffffffu, c1, o);
            c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        }
        if (lane == 0) {
            const int32_t nn = n_ref - c1, sum = 2 * c0 + c2, sumsq = 4 * c0 + c2;
            const double ni = (double)nn;
            const double d = ni * (double)sumsq - (double)sum * (double)sum;
            const double n = (double)n_ref;
            rowN[g] = nn;
            rowS[g] = sum;
            const double rr = sqrt(tau * (n - 1.0) / (n * ni * d));
            rowR[g] = rr;
            rowC[g] = make_double2((double)sum, rr);      // {S_i, r_i} as one 16-byte record: the Gram epilogues stage it with cp.async
        }
        __syncwarp();
        cur = (cur + 1 == kDecRing) ? 0 : cur + 1;
    }
}

// ------------------------------------------------------------------------------------------
// Per-block "has missing calls" flag from the counts the decoder just wrote: one warp per listed block.  The
// correlation builder reads flags[b] per tile (one integer plane or four); any[0] tells the four-plane kernel whether
// there is anything to do at all.
// ------------------------------------------------------------------------------------------
__global__ void block_flags_kernel(const BlockDesc* __restrict__ blocks, const int32_t* __restrict__ list, int32_t n_list,
                                   const int32_t* __restrict__ rowN, int32_t n_ref, int32_t* __restrict__ flags,
                                   int32_t* __restrict__ any) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n_list) return;
    const int lane = threadIdx.x & 31;
    const int b = list ? list[i] : i;
    const BlockDesc bd = blocks[b];
    int miss = 0;
    for (int j = lane; j < bd.m; j += 32) miss |= (rowN[bd.goff + j] != n_ref);
    miss = __any_sync(0xffffffffu, miss);
    if (lane == 0) {
        flags[b] = miss ? 1 : 0;
        if (miss) atomicOr(any, 1);
    }
}

// ------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------
cudaError_t launch_block_flags(const BlockDesc* blocks, const int32_t* list, int32_t n_list, const int32_t* rowN,
                               int32_t n_ref, int32_t* flags, int32_t* any, cudaStream_t st) {
    if (n_list == 0) return cudaSuccess;
    block_flags_kernel<<<(n_list + 7) / 8, 256, 0, st>>>(blocks, list, n_list, rowN, n_ref, flags, any);
    return cudaGetLastError();
}

static int stage_bytes(int32_t pitch) { return ((pitch + 15 + 16 + 15) / 16) * 16 + 16; }
static constexpr size_t kStageSmemMax = 200 * 1024;

// Largest reference panel the row-staging kernels take: one warp must hold two whole .bed rows in shared memory.
int32_t decode_max_n_ref() {
    int32_t pitch = (int32_t)(kStageSmemMax / 2) - 64;
    while (2 * (size_t)stage_bytes(pitch) > kStageSmemMax) --pitch;
    return pitch * 4;
}
// warps per CTA for `ring` staging buffers per warp (8, halved until the buffers fit; 0: even one warp does not fit)
static int stage_warps(int buf, int ring) {
    int w = kWarpsPerCta;
    while (w >= 1 && (size_t)w * ring * buf > kStageSmemMax) w >>= 1;
    return w;
}

cudaError_t launch_snp_stats(const uint8_t* bed, int64_t n_snp, int32_t n_ref, SnpStat* stats, int n_sm,
                             cudaStream_t st) {
    const int32_t pitch = (n_ref + 3) / 4;
    const int buf = stage_bytes(pitch);
    const int wpc = stage_warps(buf, 2);
    if (wpc < 1) return cudaErrorInvalidValue;
    const size_t smem = (size_t)wpc * 2 * buf;
    cudaError_t e = cudaFuncSetAttribute(snp_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t ctas = (n_snp + wpc - 1) / wpc;
    const int64_t cap = (int64_t)n_sm * 8;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    snp_stats_kernel<<<(unsigned)ctas, wpc * 32, smem, st>>>(bed, n_snp, n_ref, pitch, buf, stats);
    return cudaGetLastError();
}

template <int RING>
static cudaError_t launch_decode_t(const uint8_t* bed, int32_t n_ref, int32_t pitch, int32_t n_pad, int buf, int wpc,
                                   const uint32_t* row_src, const int32_t* row_crow, const int32_t* row_mrow, int64_t g0,
                                   int64_t n_rows, double tau, int8_t* codes, uint8_t* dirty, int32_t* rowN, int32_t* rowS,
                                   double* rowR, double2* rowC, const int32_t* gate, int n_sm, cudaStream_t st) {
    const size_t smem = (size_t)wpc * RING * buf;
    cudaError_t e = cudaFuncSetAttribute(decode_rows_kernel<RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t ctas = (n_rows + wpc - 1) / wpc;
    const int64_t cap = (int64_t)n_sm * 8;
    if (ctas > cap) ctas = cap;
    decode_rows_kernel<RING><<<(unsigned)ctas, wpc * 32, smem, st>>>(bed, n_ref, pitch, n_pad, buf, row_src, row_crow, row_mrow, g0,
                                                                   n_rows, tau, codes, dirty, rowN, rowS, rowR, rowC, gate);
    return cudaGetLastError();
}

template <int RING>
static cudaError_t launch_pack_t(const uint8_t* bed, int32_t n_ref, int32_t pitch, int32_t n_pad, int buf, int wpc,
                                 const uint32_t* row_src, int64_t g0, int64_t n_rows, double tau, uint32_t* packed,
                                 int32_t* rowN, int32_t* rowS, double* rowR, double2* rowC, int n_sm, cudaStream_t st) {
    const size_t smem = (size_t)wpc * RING * buf;
    cudaError_t e = cudaFuncSetAttribute(pack_rows_kernel<RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t ctas = (n_rows + wpc - 1) / wpc;
    const int64_t cap = (int64_t)n_sm * 8;
    if (ctas > cap) ctas = cap;
    pack_rows_kernel<RING><<<(unsigned)ctas, wpc * 32, smem, st>>>(bed, n_ref, pitch, n_pad, buf, row_src, g0, n_rows, tau, packed,
                                                                 rowN, rowS, rowR, rowC);
    return cudaGetLastError();
}
// SNP rows [g0, g0 + n_rows) of the plan -> packed 2-bit rows (pitch n_pad / 4 bytes, row g at packed + g * n_pad / 16 words)
cudaError_t launch_pack_rows(const uint8_t* bed, int32_t n_ref, int32_t n_pad, const uint32_t* row_src, int64_t g0, int64_t n_rows,
                             double tau, uint32_t* packed, int32_t* rowN, int32_t* rowS, double* rowR, double2* rowC, int n_sm, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    const int32_t pitch = (n_ref + 3) / 4;
    const int buf = stage_bytes(pitch);
    if (stage_warps(buf, 4) == kWarpsPerCta)
        return launch_pack_t<4>(bed, n_ref, pitch, n_pad, buf, kWarpsPerCta, row_src, g0, n_rows, tau, packed, rowN, rowS, rowR, rowC, n_sm, st);
    const int wpc = stage_warps(buf, 2);
    if (wpc < 1) return cudaErrorInvalidValue;
    return launch_pack_t<2>(bed, n_ref, pitch, n_pad, buf, wpc, row_src, g0, n_rows, tau, packed, rowN, rowS, rowR, rowC, n_sm, st);
}

// SNP rows [g0, g0 + n_rows) of the plan
cudaError_t launch_decode_rows(const uint8_t* bed, int32_t n_ref, int32_t n_pad, const uint32_t* row_src,
                               const int32_t* row_crow, const int32_t* row_mrow, int64_t g0, int64_t n_rows, double tau,
                               int8_t* codes, uint8_t* dirty, int32_t* rowN, int32_t* rowS, double* rowR, double2* rowC, const int32_t* gate,
                               int n_sm, cudaStream_t st) {
    if (n_rows == 0) return cudaSuccess;
    const int32_t pitch = (n_ref + 3) / 4;
    const int buf = stage_bytes(pitch);
    // four staging buffers per warp and eight warps per CTA while they fit; long rows fall back to two buffers, then
    // to fewer warps (n_ref up to decode_max_n_ref())
    if (stage_warps(buf, 4) == kWarpsPerCta)
        return launch_decode_t<4>(bed, n_ref, pitch, n_pad, buf, kWarpsPerCta, row_src, row_crow, row_mrow, g0, n_rows, tau, codes, dirty, rowN, rowS, rowR, rowC, gate, n_sm, st);
    const int wpc = stage_warps(buf, 2);
    if (wpc < 1) return cudaErrorInvalidValue;
    return launch_decode_t<2>(bed, n_ref, pitch, n_pad, buf, wpc, row_src, row_crow, row_mrow, g0, n_rows, tau, codes, dirty, rowN, rowS, rowR, rowC, gate, n_sm, st);
}

// Code row and mask row of every SNP row (the decoder's scatter maps): a function of the block layout, so they are written
// here instead of travelling with the plan blob (8 bytes per SNP less to build on the host and to push across PCIe).
__global__ void fill_rowmaps_kernel(const BlockDesc* __restrict__ blocks, int32_t* __restrict__ row_crow, int32_t* __restrict__ row_mrow) {
    const BlockDesc bd = blocks[blockIdx.x];
    for (int j = threadIdx.x; j < bd.m; j += blockDim.x) {
        row_crow[bd.goff + j] = bd.croff + j;
        row_mrow[bd.goff + j] = bd.croff + bd.m + j;
    }
}
cudaError_t launch_fill_rowmaps(const BlockDesc* blocks, int32_t n_blocks, int32_t* row_crow, int32_t* row_mrow, cudaStream_t st) {
    if (n_blocks == 0) return cudaSuccess;
    fill_rowmaps_kernel<<<n_blocks, 256, 0, st>>>(blocks, row_crow, row_mrow);
    return cudaGetLastError();
}

}  // namespace dbslmm
