// The dense intersection contraction of intersect_mma.cu on CTA PAIRS (tcgen05 cta_group::2).
//
// What bounds the one-CTA kernel (ncu, profiles/kernels_r01e.md): the tensor pipe is busy 78 % of the
// cycles and the rest is the operand expansion -- every K32 step of a 128 x 256 tile needs 384 operand rows
// of 32 bytes written to shared memory (96 B/clk of the 128 B/clk store bandwidth).  A pair of CTAs on the
// two SMs of a TPC computes a 256 x 256 tile with ONE tcgen05.mma.cta_group::2 (M256 N256 K32): each CTA
// expands only its own 128 rows of A and its own 128 columns of B (256 operand rows per step instead of
// 384, per SM), the tensor cores of both SMs read both halves of B.  Same u8 operand trick, same int32
// accumulator, same results bit for bit.
//
// Per CTA (rank 0 = leader, rank 1 = peer):
//   warps 1..8   expanders, one THREAD per mask (128 rows then 128 columns of this CTA's halves), as in
//                intersect_mma.cu; after filling a stage, one lane per warp arrives on the LEADER's "full"
//                barrier (the peer's warps through a cluster-scope remote arrive)
//   warp 0       leader only: one thread issues the MMAs of a stage once all 16 warps of the pair have
//                arrived; tcgen05.commit multicast hands the stage back to both CTAs ("empty" barriers)
//   warps 1..8   epilogue: each CTA reads its own 128 accumulator rows from its tensor memory
#include "common.cuh"
#include "async.cuh"
#include "mma_common.cuh"

#define M2_TM 256                            // rows of a pair tile (128 per CTA)
#define M2_TN 256                            // columns of a pair tile (128 per CTA as B operand rows)
#define M2_HALF 128
#define M2_SLABS 2                           // 128-pixel slabs per operand stage
#define M2_NO 3                              // operand stages
#define M2_PF 2                              // iterations of packed bits prefetched in registers
#define M2_A_BYTES (M2_HALF * 128)
#define M2_B_BYTES (M2_HALF * 128)
#define M2_SLAB_BYTES (M2_A_BYTES + M2_B_BYTES)
#define M2_OP_BYTES (M2_SLABS * M2_SLAB_BYTES)
#define M2_EXP_WARPS 8
#define M2_THREADS ((1 + M2_EXP_WARPS) * 32)
#define M2_TMEM_COLS 256

#define M2_OFF_OPS 0
#define M2_OFF_BAR (M2_NO * M2_OP_BYTES)
#define M2_N_BARS (2 * M2_NO + 1)
#define M2_OFF_MISC (M2_OFF_BAR + M2_N_BARS * 8)
#define M2_SMEM_BYTES (M2_OFF_MISC + 64 + 1024)

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(M2_THREADS, 1)
intersect_mma_pair_kernel(const MmaArgs p)
{
    extern __shared__ uint8_t smem_raw[];
    const u32 base = (smem_u32(smem_raw) + 1023u) & ~1023u;            // identical in both CTAs of the pair
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));
    u32 *s_misc = reinterpret_cast<u32 *>(gen + M2_OFF_MISC);          // [0] tmem base, [1..4] k-range reduction
    const u32 bar0 = base + M2_OFF_BAR;
    auto bar_op_full = [&](u32 s) { return bar0 + 8u * s; };           // used in the leader only
    auto bar_op_empty = [&](u32 s) { return bar0 + 8u * (M2_NO + s); };
    const u32 bar_acc = bar0 + 8u * (2 * M2_NO);

    const u32 tid = threadIdx.x, wid = tid >> 5, lane = tid & 31u;
    const u32 rank = cluster_ctarank();
    const u32 tile = blockIdx.x >> 1;
    const int g = p.tile_grp[tile], m0 = p.tile_m0[tile] + (int)(rank * M2_HALF), n0 = p.tile_n0[tile];
    const int G = p.grp_row_count[g], P = p.grp_col_count[g];
    const int rb = p.grp_row_begin[g], cb = p.grp_col_begin[g];

    if (tid == 0) {
        for (u32 s = 0; s < M2_NO; s++) { mbar_init(bar_op_full(s), 2 * M2_EXP_WARPS); mbar_init(bar_op_empty(s), 1); }
        mbar_init(bar_acc, 1);
        s_misc[1] = 0xffffffffu; s_misc[2] = 0u;      // rows: min lo, max hi
        s_misc[3] = 0xffffffffu; s_misc[4] = 0u;      // cols: min lo, max hi
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 0) {                                   // the same warp of both CTAs allocates the pair's columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(base + M2_OFF_MISC), "r"((u32)M2_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const u32 slot = tid - 32u;                      // 0..255 for warps 1..8
    const bool is_b = slot >= M2_HALF;
    u32 lo = 0, hi = 0;
    const uint4 *src = nullptr;
    if (wid > 0) {
        int mask = -1;
        if (!is_b) {
            const int pos = m0 + (int)slot;
            if (pos < G) mask = p.row_mask[rb + (p.row_order ? p.row_order[rb + pos] : pos)];
        } else {
            const int pos = n0 + (int)(rank * M2_HALF) + (int)(slot - M2_HALF);
            if (pos < P) mask = cb + (p.col_order ? p.col_order[cb + pos] : pos);
        }
        if (mask >= 0) {
            const uint2 sp = p.span[mask];
            lo = sp.x; hi = sp.y;
            src = p.bits + p.bits_off[mask] - p.reg[mask].x;
        }
        const u32 wlo = warp_min(hi > lo ? lo : 0xffffffffu), whi = warp_max(hi > lo ? hi : 0u);
        if (lane == 0 && whi > 0) {
            atomicMin(&s_misc[is_b ? 3 : 1], wlo);
            atomicMax(&s_misc[is_b ? 4 : 2], whi);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                               // barriers initialised and slab ranges written in both CTAs
    tc_fence_after();
    const u32 tmem = s_misc[0];
    // slab range of the PAIR tile: rows of both CTAs, columns of both CTAs
    const u32 peer_misc = cluster_map(base + M2_OFF_MISC, rank ^ 1u);
    const u32 rlo = min(s_misc[1], cluster_lds_u32(peer_misc + 4)), rhi = max(s_misc[2], cluster_lds_u32(peer_misc + 8));
    const u32 clo = min(s_misc[3], cluster_lds_u32(peer_misc + 12)), chi = max(s_misc[4], cluster_lds_u32(peer_misc + 16));
    const u32 klo = max(rlo, clo) & ~1u, khi = min(rhi, chi);
    const u32 nslab = khi > klo ? khi - klo : 0u;
    const u32 niter = (nslab + M2_SLABS - 1) / M2_SLABS;

    if (wid == 0) {
        if (rank == 0 && lane == 0) {
            // ---------------- MMA issuer: one thread of the leader ------------------------------------
            for (u32 it = 0; it < niter; it++) {
                const u32 o = it % M2_NO;
                mbar_wait(bar_op_full(o), (it / M2_NO) & 1u);
                tc_fence_after();
#pragma unroll
                for (u32 sl = 0; sl < M2_SLABS; sl++) {
                    const u32 a_addr = base + M2_OFF_OPS + o * M2_OP_BYTES + sl * M2_SLAB_BYTES;
                    const u64 ad = smem_desc_sw128(a_addr), bd = smem_desc_sw128(a_addr + M2_A_BYTES);
#pragma unroll
                    for (u32 k = 0; k < 4; k++)
                        tc_mma_i8_pair(tmem, ad + 2u * k, bd + 2u * k, mma_idesc(M2_TM, M2_TN), (it | sl | k) ? 1u : 0u);
                }
                tc_commit_pair(bar_op_empty(o));
            }
            if (niter) tc_commit_pair(bar_acc);
        }
        __syncwarp();
    } else {
        // ---------------- expanders ---------------------------------------------------------------------
        const u32 r = is_b ? slot - M2_HALF : slot, r7 = r & 7u;
        const u32 row_off = (is_b ? (u32)M2_A_BYTES : 0u) + (r >> 3) * 1024u + r7 * 128u;
        const u32 full0 = cluster_map(bar_op_full(0), 0u);               // the leader's "full" barriers
        uint4 buf[M2_PF + 1][M2_SLABS];
        const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
        auto fetch = [&](u32 it, uint4 (&d)[M2_SLABS]) {
#pragma unroll
            for (u32 sl = 0; sl < M2_SLABS; sl++) {
                const u32 k = klo + it * M2_SLABS + sl;
                d[sl] = (k >= lo && k < hi) ? ldg_v4(src + k) : zero4;
            }
        };
#pragma unroll
        for (u32 f = 0; f < M2_PF; f++) {
            if (f < niter) fetch(f, buf[f]);
        }
        for (u32 it = 0; it < niter; it++) {
            if (it + M2_PF < niter) fetch(it + M2_PF, buf[M2_PF]);
            const u32 o = it % M2_NO;
            mbar_wait(bar_op_empty(o), ((it / M2_NO) & 1u) ^ 1u);      // local barrier, signalled by tcgen05.commit
            const u32 stage = base + M2_OFF_OPS + o * M2_OP_BYTES + row_off;
#pragma unroll
            for (u32 sl = 0; sl < M2_SLABS; sl++) {
                if (is_b) expand_chunk<true>(buf[0][sl], stage + sl * M2_SLAB_BYTES, r7);
                else expand_chunk<false>(buf[0][sl], stage + sl * M2_SLAB_BYTES, r7);
            }
            fence_proxy_async();                                 // generic-proxy stores -> visible to the MMA
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) mbar_arrive(bar_op_full(o));          // leader: its own barrier
                else mbar_arrive_cluster(full0 + 8u * o);            // peer: remote arrive
            }
#pragma unroll
            for (u32 f = 0; f < M2_PF; f++) {
#pragma unroll
                for (u32 sl = 0; sl < M2_SLABS; sl++) buf[f][sl] = buf[f + 1][sl];
            }
        }
        // ---------------- epilogue: this CTA's 128 rows, TMEM -> int32 matrix ---------------------------
        const i64 off = p.grp_imat_off[g];
        const u32 quarter = wid & 3u, half = (wid - 1u) >> 2;
        const int rpos = m0 + (int)(32u * quarter + lane);
        const int row = (rpos < G && p.row_order) ? p.row_order[rb + rpos] : rpos;
        if (niter) {
            mbar_wait(bar_acc, 0u);
            tc_fence_after();
        }
#pragma unroll 1
        for (u32 cbk = 0; cbk < 4; cbk++) {
            u32 v[32];
            const u32 col0 = half * 128u + cbk * 32u;
            if (niter) {
                tc_ld32(tmem + ((32u * quarter) << 16) + col0, v);
            } else {
#pragma unroll
                for (int t = 0; t < 32; t++) v[t] = 0u;
            }
            if (rpos < G) {
                int *orow = p.imat + off + (i64)row * P;
#pragma unroll
                for (int t = 0; t < 32; t++) {
                    const int cpos = n0 + (int)col0 + t;
                    if (cpos < P) orow[p.col_order ? p.col_order[cb + cpos] : cpos] = (int)(v[t] >> 7);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    cluster_sync_all();                               // nobody leaves while the other CTA can still signal it
    if (wid == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((u32)M2_TMEM_COLS) : "memory");
    }
}

extern "C" int ampis_mma_pair_tile_rows(void) { return M2_TM; }
extern "C" int ampis_mma_pair_tile_cols(void) { return M2_TN; }

extern "C" int ampis_intersect_tcgen05_pair(const void *d_bits, const int64_t *d_bits_off, const uint32_t *d_reg,
                                            const uint32_t *d_span, const int32_t *d_row_mask,
                                            const int32_t *d_row_order, const int32_t *d_col_order,
                                            const int32_t *d_tile_grp, const int32_t *d_tile_m0,
                                            const int32_t *d_tile_n0, int32_t n_tiles,
                                            const int32_t *d_grp_row_begin, const int32_t *d_grp_row_count,
                                            const int32_t *d_grp_col_begin, const int32_t *d_grp_col_count,
                                            const int64_t *d_grp_imat_off, int32_t *d_imat, void *stream)
{
    AMPIS_REQUIRE(n_tiles >= 0, "n_tiles < 0");
    if (n_tiles == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits && d_bits_off && d_reg && d_span && d_row_mask && d_tile_grp && d_tile_m0 && d_tile_n0 &&
                      d_grp_row_begin && d_grp_row_count && d_grp_col_begin && d_grp_col_count && d_grp_imat_off &&
                      d_imat, "null pointer");
    cudaError_t e = cudaFuncSetAttribute(intersect_mma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         M2_SMEM_BYTES);
    if (e != cudaSuccess) { ampis_set_error("intersect_mma_pair_kernel smem: %s", cudaGetErrorString(e)); return AMPIS_ECUDA; }
    MmaArgs a;
    a.bits = (const uint4 *)d_bits; a.bits_off = d_bits_off; a.reg = (const uint2 *)d_reg;
    a.span = (const uint2 *)d_span; a.row_mask = d_row_mask;
    a.row_order = d_row_order; a.col_order = d_col_order;
    a.tile_grp = d_tile_grp; a.tile_m0 = d_tile_m0; a.tile_n0 = d_tile_n0;
    a.grp_row_begin = d_grp_row_begin; a.grp_row_count = d_grp_row_count;
    a.grp_col_begin = d_grp_col_begin; a.grp_col_count = d_grp_col_count;
    a.grp_imat_off = d_grp_imat_off; a.imat = d_imat;
    intersect_mma_pair_kernel<<<2 * n_tiles, M2_THREADS, M2_SMEM_BYTES, as_stream(stream)>>>(a);   // clusters of 2
    AMPIS_CHECK_LAUNCH("intersect_mma_pair_kernel");
    return AMPIS_OK;
}
