// Dense intersection matrices on the 5th-generation tensor cores (tcgen05, sm_100a only).
//
// For images whose instances overlap heavily (candidate density of tens of percent) the
// bbox-culled AND+popc walk of intersect.cu re-reads every mask once per overlapping partner.
// There the whole G x P matrix of an image is cheaper as ONE integer contraction
//
//        I[r][c] = sum_k A[r][k] * B[c][k]          k = pixel index, A, B in {0,1}
//
// i.e. exactly the quantity pycocotools' rleIou run walk accumulates for every pair
// (analyze.py:108,158 via RLE.iou; powder.py:82 via RLE.merge(intersect=True) + RLE.area), with
// no pruning at all.  int32 accumulation makes it bit-exact.
//
// One CTA computes one 128 x 256 tile of one image's matrix:
//   warps 1..12  expanders, one THREAD per mask of the tile (128 rows + 256 columns): the thread
//                streams its mask's packed bits from HBM/L2 (two 128-bit chunks = one 32-byte sector
//                per iteration, prefetched two iterations ahead in registers) and expands them to a
//                u8 operand row in shared memory, K-major with the 128-byte swizzle tcgen05 expects
//                (one 128-pixel slab = one 128-byte row; a stage holds two slabs)
//   warp 0       one thread issues tcgen05.mma.kind::i8 (M128 N256 K32, 8 per stage) with the
//                accumulator in tensor memory (256 columns); tcgen05.commit hands stages back
//   warps 1..8   epilogue: tcgen05.ld of the accumulator, >> 7, int32 stores
//
// Bit expansion without shifts: a byte of packed bits is replicated by one PRMT and masked, so
// pixel j of an A row becomes the value 2^(j%8) (bit kept in place) and pixel j of a B row
// becomes 2^(7-j%8) (bit kept in place of the bit-reversed word).  Every product of two set
// pixels is then exactly 128, every other product 0, and I = accumulator >> 7.  u8 x u8 -> s32
// wraps only above 2^32/128 = 33.5 M pixels per mask (5792 x 5792).
#include "common.cuh"
#include "async.cuh"
#include "mma_common.cuh"

#define MMA_TM 128
#define MMA_TN 256
#define MMA_SLOTS (MMA_TM + MMA_TN)         // mask slots of a tile: 128 rows then 256 columns
#define MMA_SLABS 2                          // 128-pixel slabs per operand stage
#define MMA_NO 2                             // operand stages
#define MMA_PF 2                             // iterations of packed bits prefetched in registers
#define MMA_A_BYTES (MMA_TM * 128)
#define MMA_B_BYTES (MMA_TN * 128)
#define MMA_SLAB_BYTES (MMA_A_BYTES + MMA_B_BYTES)
#define MMA_OP_BYTES (MMA_SLABS * MMA_SLAB_BYTES)
#define MMA_EXP_WARPS (MMA_SLOTS / 32)
#define MMA_THREADS ((1 + MMA_EXP_WARPS) * 32)
#define MMA_TMEM_COLS 256

// dynamic shared memory map (base aligned to 1024 by hand)
#define MMA_OFF_OPS 0
#define MMA_OFF_BAR (MMA_NO * MMA_OP_BYTES)
#define MMA_N_BARS (2 * MMA_NO + 1)
#define MMA_OFF_MISC (MMA_OFF_BAR + MMA_N_BARS * 8)
#define MMA_SMEM_BYTES (MMA_OFF_MISC + 64 + 1024)

__global__ void __launch_bounds__(MMA_THREADS, 1)
intersect_mma_kernel(const MmaArgs p)
{
    extern __shared__ uint8_t smem_raw[];
    const u32 base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen = smem_raw + (base - smem_u32(smem_raw));            // generic pointer to the same place
    u32 *s_misc = reinterpret_cast<u32 *>(gen + MMA_OFF_MISC);        // [0] tmem base, [1..4] k-range reduction
    const u32 bar0 = base + MMA_OFF_BAR;
    auto bar_op_full = [&](u32 s) { return bar0 + 8u * s; };
    auto bar_op_empty = [&](u32 s) { return bar0 + 8u * (MMA_NO + s); };
    const u32 bar_acc = bar0 + 8u * (2 * MMA_NO);

    const u32 tid = threadIdx.x, wid = tid >> 5, lane = tid & 31u;
    const int g = p.tile_grp[blockIdx.x], m0 = p.tile_m0[blockIdx.x], n0 = p.tile_n0[blockIdx.x];
    const int G = p.grp_row_count[g], P = p.grp_col_count[g];
    const int rb = p.grp_row_begin[g], cb = p.grp_col_begin[g];

    if (tid == 0) {
        for (u32 s = 0; s < MMA_NO; s++) { mbar_init(bar_op_full(s), MMA_EXP_WARPS); mbar_init(bar_op_empty(s), 1); }
        mbar_init(bar_acc, 1);
        s_misc[1] = 0xffffffffu; s_misc[2] = 0u;      // rows: min lo, max hi
        s_misc[3] = 0xffffffffu; s_misc[4] = 0u;      // cols: min lo, max hi
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(base + MMA_OFF_MISC), "r"((u32)MMA_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    // this thread's mask (expander threads): where its packed bits live, which slabs hold 1-pixels
    const u32 slot = tid - 32u;                      // 0..383 for warps 1..12
    const bool is_b = slot >= MMA_TM;
    u32 lo = 0, hi = 0;
    const uint4 *src = nullptr;
    if (wid > 0) {
        // tiles are cut from the SORTED row / column lists (by first occupied slab) when orders are given:
        // masks of a tile are then neighbours in the image and the slab range of the tile shrinks
        int mask = -1;
        if (!is_b) {
            const int pos = m0 + (int)slot;
            if (pos < G) mask = p.row_mask[rb + (p.row_order ? p.row_order[rb + pos] : pos)];
        } else {
            const int pos = n0 + (int)(slot - MMA_TM);
            if (pos < P) mask = cb + (p.col_order ? p.col_order[cb + pos] : pos);
        }
        if (mask >= 0) {
            const uint2 sp = p.span[mask];
            lo = sp.x; hi = sp.y;
            src = p.bits + p.bits_off[mask] - p.reg[mask].x;        // chunk k of the mask is src[k]
        }
        // warp-level then CTA-level range of occupied slabs, rows and columns separately
        const u32 wlo = warp_min(hi > lo ? lo : 0xffffffffu), whi = warp_max(hi > lo ? hi : 0u);
        if (lane == 0 && whi > 0) {
            atomicMin(&s_misc[is_b ? 3 : 1], wlo);
            atomicMax(&s_misc[is_b ? 4 : 2], whi);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const u32 tmem = s_misc[0];
    // slabs in which a row AND a column have pixels; start on an even chunk (32-byte sectors)
    const u32 klo = max(s_misc[1], s_misc[3]) & ~1u, khi = min(s_misc[2], s_misc[4]);
    const u32 nslab = khi > klo ? khi - klo : 0u;
    const u32 niter = (nslab + MMA_SLABS - 1) / MMA_SLABS;

    if (wid == 0) {
        // ---------------- MMA issuer: one thread ------------------------------------------------------
        if (lane == 0) {
            for (u32 it = 0; it < niter; it++) {
                const u32 o = it % MMA_NO;
                mbar_wait(bar_op_full(o), (it / MMA_NO) & 1u);
                tc_fence_after();
#pragma unroll
                for (u32 sl = 0; sl < MMA_SLABS; sl++) {
                    const u32 a_addr = base + MMA_OFF_OPS + o * MMA_OP_BYTES + sl * MMA_SLAB_BYTES;
                    const u64 ad = smem_desc_sw128(a_addr), bd = smem_desc_sw128(a_addr + MMA_A_BYTES);
#pragma unroll
                    for (u32 k = 0; k < 4; k++)    // 4 x K32 inside the 128-byte swizzle atom: +32 bytes each
                        tc_mma_i8(tmem, ad + 2u * k, bd + 2u * k, mma_idesc(MMA_TM, MMA_TN), (it | sl | k) ? 1u : 0u);
                }
                tc_commit(bar_op_empty(o));        // implies tcgen05.fence::before_thread_sync
            }
            if (niter) tc_commit(bar_acc);
        }
        __syncwarp();
    } else {
        // ---------------- expanders: bits -> u8 operand rows -----------------------------------------
        const u32 r = is_b ? slot - MMA_TM : slot, r7 = r & 7u;
        const u32 row_off = (is_b ? (u32)MMA_A_BYTES : 0u) + (r >> 3) * 1024u + r7 * 128u;
        uint4 buf[MMA_PF + 1][MMA_SLABS];
        const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
        auto fetch = [&](u32 it, uint4 (&d)[MMA_SLABS]) {
#pragma unroll
            for (u32 sl = 0; sl < MMA_SLABS; sl++) {
                const u32 k = klo + it * MMA_SLABS + sl;
                d[sl] = (k >= lo && k < hi) ? ldg_v4(src + k) : zero4;
            }
        };
#pragma unroll
        for (u32 f = 0; f < MMA_PF; f++) {
            if (f < niter) fetch(f, buf[f]);
        }
        for (u32 it = 0; it < niter; it++) {
            if (it + MMA_PF < niter) fetch(it + MMA_PF, buf[MMA_PF]);
            const u32 o = it % MMA_NO;
            mbar_wait(bar_op_empty(o), ((it / MMA_NO) & 1u) ^ 1u);
            const u32 stage = base + MMA_OFF_OPS + o * MMA_OP_BYTES + row_off;
#pragma unroll
            for (u32 sl = 0; sl < MMA_SLABS; sl++) {
                if (is_b) expand_chunk<true>(buf[0][sl], stage + sl * MMA_SLAB_BYTES, r7);
                else expand_chunk<false>(buf[0][sl], stage + sl * MMA_SLAB_BYTES, r7);
            }
            fence_proxy_async();                                 // generic-proxy stores -> visible to the MMA
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_op_full(o));
#pragma unroll
            for (u32 f = 0; f < MMA_PF; f++) {
#pragma unroll
                for (u32 sl = 0; sl < MMA_SLABS; sl++) buf[f][sl] = buf[f + 1][sl];
            }
        }
        // ---------------- epilogue: TMEM -> int32 matrix (warps 1..8) --------------------------------
        if (wid <= 8) {
            const i64 off = p.grp_imat_off[g];
            const u32 quarter = wid & 3u, half = (wid - 1u) >> 2;   // TMEM lanes 32*quarter.., columns 128*half..
            const int rpos = m0 + (int)(32u * quarter + lane);
            const int row = (rpos < G && p.row_order) ? p.row_order[rb + rpos] : rpos;      // unsorted index for the output
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
    }
    __syncthreads();
    if (wid == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((u32)MMA_TMEM_COLS) : "memory");
    }
}

extern "C" int ampis_mma_tile_rows(void) { return MMA_TM; }
extern "C" int ampis_mma_tile_cols(void) { return MMA_TN; }

extern "C" int ampis_intersect_tcgen05(const void *d_bits, const int64_t *d_bits_off, const uint32_t *d_reg,
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
    // per launch: the attribute is per device, and a process may drive more than one
    cudaError_t e = cudaFuncSetAttribute(intersect_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         MMA_SMEM_BYTES);
    if (e != cudaSuccess) { ampis_set_error("intersect_mma_kernel smem: %s", cudaGetErrorString(e)); return AMPIS_ECUDA; }
    MmaArgs a;
    a.bits = (const uint4 *)d_bits; a.bits_off = d_bits_off; a.reg = (const uint2 *)d_reg;
    a.span = (const uint2 *)d_span; a.row_mask = d_row_mask;
    a.row_order = d_row_order; a.col_order = d_col_order;
    a.tile_grp = d_tile_grp; a.tile_m0 = d_tile_m0; a.tile_n0 = d_tile_n0;
    a.grp_row_begin = d_grp_row_begin; a.grp_row_count = d_grp_row_count;
    a.grp_col_begin = d_grp_col_begin; a.grp_col_count = d_grp_col_count;
    a.grp_imat_off = d_grp_imat_off; a.imat = d_imat;
    intersect_mma_kernel<<<n_tiles, MMA_THREADS, MMA_SMEM_BYTES, as_stream(stream)>>>(a);
    AMPIS_CHECK_LAUNCH("intersect_mma_kernel");
    return AMPIS_OK;
}

// ---------------------------------------------------------------------------------------------
// Row results from a dense intersection matrix: score, first arg-max (same rules as
// intersect_rows_kernel; analyze.py:158-164, powder.py:82-86).  Warp per row.
// ---------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(256)
rows_from_imat_kernel(const int *__restrict__ imat, const i64 *__restrict__ grp_imat_off,
                      const u32 *__restrict__ area, const int *__restrict__ row_mask,
                      const int *__restrict__ row_grp, int n_rows, const int *__restrict__ grp_row_begin,
                      const int *__restrict__ grp_col_begin, const int *__restrict__ grp_col_count,
                      int *__restrict__ best_col, u32 *__restrict__ best_inter, double *__restrict__ best_score)
{
    const int r = (int)((blockIdx.x * (u32)blockDim.x + threadIdx.x) >> 5);
    if (r >= n_rows) return;
    const u32 lane = lane_id();
    const int g = row_grp[r];
    const int P = grp_col_count[g], cb = grp_col_begin[g];
    const int *irow = imat + grp_imat_off[g] + (i64)(r - grp_row_begin[g]) * P;
    const u32 ra = area[row_mask[r]];
    double best_s = 0.0;
    u32 best_i = 0;
    int best_c = MODE == AMPIS_MODE_IOU ? -1 : (P > 0 ? 0 : -1);
    for (int c = (int)lane; c < P; c += 32) {
        const u32 inter = (u32)irow[c];
        if (MODE == AMPIS_MODE_IOU) {
            const double s = inter ? (double)inter / (double)(ra + area[cb + c] - inter) : 0.0;
            if (s > best_s) { best_s = s; best_i = inter; best_c = c; }
        } else {
            if (inter > best_i) { best_i = inter; best_c = c; }
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        const double os = __shfl_xor_sync(0xffffffffu, best_s, d);
        const u32 oi = __shfl_xor_sync(0xffffffffu, best_i, d);
        const int oc = __shfl_xor_sync(0xffffffffu, best_c, d);
        bool take;
        if (MODE == AMPIS_MODE_IOU) take = os > best_s || (os == best_s && (unsigned)oc < (unsigned)best_c);
        else take = oi > best_i || (oi == best_i && (unsigned)oc < (unsigned)best_c);
        if (take) { best_s = os; best_i = oi; best_c = oc; }
    }
    if (lane == 0) {
        if (MODE == AMPIS_MODE_SAT) best_s = (double)best_i / (double)ra;
        best_col[r] = best_c;
        best_inter[r] = best_i;
        best_score[r] = best_s;
    }
}

extern "C" int ampis_rows_from_imat(const int32_t *d_imat, const int64_t *d_grp_imat_off, const uint32_t *d_area,
                                    const int32_t *d_row_mask, const int32_t *d_row_grp, int32_t n_rows,
                                    const int32_t *d_grp_row_begin, const int32_t *d_grp_col_begin,
                                    const int32_t *d_grp_col_count, int32_t mode, int32_t *d_best_col,
                                    uint32_t *d_best_inter, double *d_best_score, void *stream)
{
    AMPIS_REQUIRE(n_rows >= 0, "n_rows < 0");
    AMPIS_REQUIRE(mode == AMPIS_MODE_IOU || mode == AMPIS_MODE_SAT, "bad mode");
    if (n_rows == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_imat && d_grp_imat_off && d_area && d_row_mask && d_row_grp && d_grp_row_begin &&
                      d_grp_col_begin && d_grp_col_count && d_best_col && d_best_inter && d_best_score,
                  "null pointer");
    const unsigned blocks = (unsigned)(((i64)n_rows * 32 + 255) / 256);
    if (mode == AMPIS_MODE_IOU)
        rows_from_imat_kernel<AMPIS_MODE_IOU><<<blocks, 256, 0, as_stream(stream)>>>(
            d_imat, d_grp_imat_off, d_area, d_row_mask, d_row_grp, n_rows, d_grp_row_begin, d_grp_col_begin,
            d_grp_col_count, d_best_col, d_best_inter, d_best_score);
    else
        rows_from_imat_kernel<AMPIS_MODE_SAT><<<blocks, 256, 0, as_stream(stream)>>>(
            d_imat, d_grp_imat_off, d_area, d_row_mask, d_row_grp, n_rows, d_grp_row_begin, d_grp_col_begin,
            d_grp_col_count, d_best_col, d_best_inter, d_best_score);
    AMPIS_CHECK_LAUNCH("rows_from_imat_kernel");
    return AMPIS_OK;
}
