// common.cuh -- shared device helpers (sm_100a PTX wrappers) and plan structures.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dbslmm {

// ----------------------------------------------------------------------------------------
// Plan structures shared by host and device
// ----------------------------------------------------------------------------------------
struct SnpStat {            // per .bed row, produced by snp_stats_kernel at load time
    int32_t n_nonmiss;      // individuals with a called genotype
    int32_t sum;            // sum of allele counts g in {0,1,2} over called genotypes
    int32_t sumsq;          // sum of g^2
    int32_t pad;
};

struct BlockDesc {
    int64_t moff;           // offset (doubles) of this block's matrix inside sigma / L buffers
    int32_t goff;           // first SNP row (block-ordered global index: small SNPs then large)
    int32_t croff;          // first int8 code row; mask rows (if any) follow the m genotype rows
    int32_t m;              // m_s + m_l
    int32_t ms;             // m_s
    int32_t mp;             // m rounded up to a multiple of 8 (padding rows are identity)
    int32_t ld;             // leading dimension in doubles (== mp); matrix has mp + 8 rows, row mp = z
    int32_t has_missing;    // 1 => four integer Gram planes (Q, A, B, N)
    int32_t out_s;          // offset of this block's small betas in the output array
    int32_t out_l;          // offset of its large betas
    int32_t nrows;          // matrix rows: mp + 8 (z row group) + appended test-genotype rows (variance side channel)
};

struct GramTile {           // one 128x128 output tile of a block's lower triangle (256 x 256 in the CTA-pair list)
    int32_t blk;
    int32_t ti;             // tile row   (rows  ti*128 .. of the block)  -> "I" side
    int32_t tj;             // tile col   (cols  tj*128 ..)               -> "J" side, tj <= ti
    int32_t pad;
};

// One tile of the one-plane Gram kernels with everything the kernel needs about its block (48 bytes = three 16-byte loads):
// a role fetches the record of tile n+1 while it works on tile n, so no dependent tile -> block -> constants chain of
// global loads sits between two tiles.
struct TileRec {
    int32_t blk, ti, tj, croff;
    int32_t goff, m, mp, ld;
    int64_t moff;
    int64_t pad;
};

// ----------------------------------------------------------------------------------------
// PTX wrappers
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}

// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); bytes % 16 == 0,
// both addresses 16-byte aligned.  Completion is signalled on `bar` (complete_tx).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// 2-D tiled TMA load (SASS UTMALDG).
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int32_t x, int32_t y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
            "r"(smem_u32(dst_smem)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y)
        : "memory");
}
// the same with an L2 eviction-priority hint (0x14F0000000000000 = evict last, 0x12F0000000000000 = evict first)
__device__ __forceinline__ void tma_load_2d_hint(void* dst_smem, const void* tmap, int32_t x, int32_t y, uint64_t* bar, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::
            "r"(smem_u32(dst_smem)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(x), "r"(y), "l"(hint)
        : "memory");
}
// 4-D tiled TMA load (SASS UTMALDG): the FP64 panel operands of the block solver (chol.cu) -- the two middle
// dimensions permute the rows of every 8-row group on their way into shared memory.
__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const void* tmap, int32_t c0, int32_t c1, int32_t c2,
                                            int32_t c3, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
            "r"(dst_smem),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// a tensor map that lives in global memory (written by the host before the launch): acquire it for the TMA unit
__device__ __forceinline__ void tensormap_acquire(const void* tmap) {
    asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}
// generic-proxy accesses (before) are ordered with async-proxy (TMA) accesses (after)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// pull `bytes` (multiple of 16, 16-byte aligned address) into L2 without a destination
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
// 128-bit shared-memory accesses by 32-bit shared address
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f64x2(uint32_t addr, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(addr), "d"(x), "d"(y) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// cp.async 16 B with zero-fill when !valid
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool valid) {
    const uint32_t n = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(n)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// FP64 tensor-core tile: D(8x8) += A(8x4, row) * B(4x8, col)   (SASS DMMA.8x8x4)
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// One lane of a fully active, converged warp (the same lane every time for the same mask).  The TMA-producer and MMA-issuer
// loops run with the WHOLE warp and guard only the issuing instructions with this: inside an `if (lane == 0)` region the
// compiler cannot prove that the operands of UTMALDG / UTCIMMA / UTCBAR (uniform registers) are warp-uniform and wraps
// every one of them in a vote / elect / broadcast loop (~18 SASS instructions each).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// Exact u32 -> f64 on the FP64 pipe: 2^52 + u as a bit pattern, minus 2^52 (one DADD).  The I2F.F64 the compiler emits for
// (double)int runs on the XU pipe at ~2 results per clock and SM on sm_100a (measured: 77 % XU-busy Gram epilogues, ncu
// round 2) -- a 128 x 128 tile of conversions then costs more than its 64 int8 MMAs.
__device__ __forceinline__ double u32_to_f64(uint32_t u) {
    return __hiloint2double(0x43300000, (int)u) - 4503599627370496.0;
}

// ---- tcgen05 / TMEM ---------------------------------------------------------------------
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)),
                 "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {      // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], int8 operands, s32 accumulate (SASS UTCIMMA).
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// ---- CTA pairs (tcgen05 cta_group::2): two CTAs of a 2-CTA cluster share one MMA of M = 256 ------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot_smem) {   // whole warp, the same warp of BOTH CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)),
                 "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {      // whole warp, the same warp of BOTH CTAs
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// 2-D tiled TMA load into THIS CTA's shared memory whose completion bytes are counted on the LEADER CTA's mbarrier
// (the barrier at the same offset in cluster rank 0: the peer bit, bit 24 of a shared-window address, cleared).
__device__ __forceinline__ void tma_load_2d_pair(void* dst_smem, const void* tmap, int32_t x, int32_t y, uint64_t* bar, uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::
            "r"(smem_u32(dst_smem)),
        "l"(tmap), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(x), "r"(y), "l"(hint)
        : "memory");
}
// D[tmem of both CTAs] (+)= A * B over the CTA pair: M = 256 (128 rows of A per CTA), B = N/2 rows per CTA.  Leader only.
__device__ __forceinline__ void umma_i8_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this offset in BOTH CTAs when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
// arrive on the mbarrier at this offset in cluster rank `rank` (no cluster-scope release: what it orders here is a TMEM
// read, fenced by tcgen05.fence::before_thread_sync; a .release.cluster would wait for every global store in flight)
__device__ __forceinline__ void mbar_arrive_rank(uint64_t* bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)),
        "r"(rank)
        : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive 32-bit columns (SASS LDTM)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory operand descriptor (canonical layout written by a
// TMA box of 128 bytes x rows with CU_TENSOR_MAP_SWIZZLE_128B): 8-row groups 1024 B apart.
// Field layout follows cute::UMMA::SmemDescriptor (start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) with SWIZZLE_128B = 2).
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;            // LBO (unused for swizzled K-major), canonical value 1
    d |= (uint64_t)(1024 >> 4) << 32;  // SBO: 8 rows * 128 B
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}

// Instruction descriptor for kind::i8, u8 x u8 -> s32, both operands K-major, M x N tile
// (cute::UMMA::InstrDescriptor: c_format [4,6)=2 (S32), a/b_format [7,10)/[10,13)=0 (U8),
//  n_dim [17,23)=N>>3, m_dim [24,29)=M>>4).
__host__ __device__ constexpr uint32_t make_i8_idesc(uint32_t M, uint32_t N) {
    return (2u << 4) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

}  // namespace dbslmm
