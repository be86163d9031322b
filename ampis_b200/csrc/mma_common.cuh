// tcgen05 / TMEM helpers and the bit -> u8 operand expansion shared by the tensor-core contraction
// kernels (intersect_mma.cu: one CTA per tile; intersect_mma2.cu: a CTA pair per tile).
#pragma once
#include "common.cuh"
#include "async.cuh"

struct MmaArgs {
    const uint4 *bits;
    const i64 *bits_off;
    const uint2 *reg;
    const uint2 *span;
    const int *row_mask;
    const int *row_order;      // optional: position inside the group's sorted row list -> row index in the group
    const int *col_order;      // optional, indexed by grp_col_begin + sorted position -> column index in the group
    const int *tile_grp, *tile_m0, *tile_n0;
    const int *grp_row_begin, *grp_row_count, *grp_col_begin, *grp_col_count;
    const i64 *grp_imat_off;
    int *imat;
};

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(u32 bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, u8 x u8 -> s32, M128 N256 K32
__device__ __forceinline__ void tc_mma_i8(u32 tmem_d, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ u64 smem_desc_sw128(u32 addr)
{
    return (u64)((addr >> 4) & 0x3fffu) | (1ull << 16) | ((u64)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: C = s32, A = B = u8, both K-major, N columns, M rows (M = 256 for a CTA pair)
__host__ __device__ constexpr u32 mma_idesc(u32 M, u32 N)
{
    return (2u << 4) | (0u << 7) | (0u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

__device__ __forceinline__ void tc_ld32(u32 taddr, u32 (&v)[32])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                   "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                   "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                   "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 32 packed pixels -> 32 operand bytes (two 16-byte chunks)
template <bool IS_B>
__device__ __forceinline__ void expand_word(u32 x, uint4 &o0, uint4 &o1)
{
    if (IS_B) x = __brev(x);
    const u32 v0 = __byte_perm(x, 0u, IS_B ? 0x3333u : 0x0000u);    // packed byte 0 in all four lanes
    const u32 v1 = __byte_perm(x, 0u, IS_B ? 0x2222u : 0x1111u);
    const u32 v2 = __byte_perm(x, 0u, IS_B ? 0x1111u : 0x2222u);
    const u32 v3 = __byte_perm(x, 0u, IS_B ? 0x0000u : 0x3333u);
    const u32 mlo = IS_B ? 0x10204080u : 0x08040201u;               // pixels 0..3 of the byte
    const u32 mhi = IS_B ? 0x01020408u : 0x80402010u;               // pixels 4..7 of the byte
    o0 = make_uint4(v0 & mlo, v0 & mhi, v1 & mlo, v1 & mhi);
    o1 = make_uint4(v2 & mlo, v2 & mhi, v3 & mlo, v3 & mhi);
}

__device__ __forceinline__ void sts_v4(u32 addr, uint4 v)
{
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ u32 lds_u32(u32 addr)
{
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

__device__ __forceinline__ uint4 ldg_v4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

// one 128-pixel chunk -> one 128-byte operand row (8 swizzled 16-byte stores)
template <bool IS_B>
__device__ __forceinline__ void expand_chunk(uint4 c, u32 row_addr, u32 r7)
{
    uint4 o0, o1;
    expand_word<IS_B>(c.x, o0, o1);
    sts_v4(row_addr + ((0u ^ r7) << 4), o0); sts_v4(row_addr + ((1u ^ r7) << 4), o1);
    expand_word<IS_B>(c.y, o0, o1);
    sts_v4(row_addr + ((2u ^ r7) << 4), o0); sts_v4(row_addr + ((3u ^ r7) << 4), o1);
    expand_word<IS_B>(c.z, o0, o1);
    sts_v4(row_addr + ((4u ^ r7) << 4), o0); sts_v4(row_addr + ((5u ^ r7) << 4), o1);
    expand_word<IS_B>(c.w, o0, o1);
    sts_v4(row_addr + ((6u ^ r7) << 4), o0); sts_v4(row_addr + ((7u ^ r7) << 4), o1);
}


// ---- CTA-pair (cta_group::2) and cluster helpers -------------------------------------------------
__device__ __forceinline__ u32 cluster_ctarank()
{
    u32 r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared::cta address) in the CTA of rank `rank`
__device__ __forceinline__ u32 cluster_map(u32 local, u32 rank)
{
    u32 r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ u32 cluster_lds_u32(u32 cluster_addr)
{
    u32 v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr) : "memory");
    return v;
}
// arrive on a barrier of another CTA of the cluster.  Plain form (release at CTA scope): a cluster-scope release
// compiles to MEMBAR.ALL.GPU + ERRBAR per arrive (measured: 19 % of the pair kernel's stall samples, stage
// handshake ~0.8 us).  The operand bytes the arrive publishes were written to this CTA's OWN shared memory and
// made visible to the async proxy by fence.proxy.async before it -- the same pattern as CUTLASS's 2-SM
// pipelines (umma_arrive_2x1SM_sm0 / ClusterBarrier::arrive(cta_id)).
__device__ __forceinline__ void mbar_arrive_cluster(u32 cluster_addr)
{
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T over the pair: M256 (128 rows per CTA) N256 (128 columns of B per CTA) K32
__device__ __forceinline__ void tc_mma_i8_pair(u32 tmem_d, u64 adesc, u64 bdesc, u32 idesc, u32 accumulate)
{
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once the pair's earlier MMAs have completed) on the barrier at the same offset in both CTAs
__device__ __forceinline__ void tc_commit_pair(u32 bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((unsigned short)3) : "memory");
}
