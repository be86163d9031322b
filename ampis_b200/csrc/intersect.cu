// The fused hot kernel: bbox-pruned mask intersections, scores and per-row arg-max.
//
// Reference behaviour being replaced:
//   analyze._piecewise_rle_match  (analyze.py:149-164)  for each GT: RLE.iou against all
//       predictions in chunks of 80 -> pycocotools rleIou = rleToBbox + bbIou pre-pass, run
//       walk only where the boxes overlap, iou = i/u with u=1 when i==0; arg-max with
//       first-max tie break and strict '>' against a running max that starts at 0.
//   powder._rle_satellite_match   (powder.py:80-86)     for each satellite:
//       area(merge(sat, particle, intersect)) / area(sat) against all particles, np.argmax.
//
// A CTA owns up to 8 consecutive rows of ONE group (image); a warp owns one row.
//   * The column masks' metadata (tight boxes, spans, stored regions, areas, arena offsets: five
//     contiguous arrays) is staged into shared memory by TMA bulk copies (cp.async.bulk + mbarrier
//     complete_tx) issued by one thread -- no per-element address arithmetic, no register staging.
//   * A warp scans all columns of the tile against its row out of shared memory and collects the
//     candidates (boxes overlap AND spans overlap; pruning cannot change a result: disjoint boxes =>
//     intersection 0 => score 0) in a list.
//   * Candidates are then intersected four at a time, eight lanes per candidate: 128-bit loads of
//     both packed masks over the overlap of their spans, AND + popc, 8-lane reduction -- so the
//     loads of four candidates are in flight together instead of one exposed round trip each
//     (61 % of the stall samples of the first version, ncu stall sampling, round 1).  Long overlaps
//     take the whole warp.
//   * The dense row of the intersection matrix is zero-filled with coalesced stores during the
//     scan; candidates patch their cells.
#include "common.cuh"
#include "async.cuh"

#define ROWS_PER_CTA AMPIS_ROWS_PER_CTA
#define COL_TILE 512
#define CAND_LIST 64
#define LONG_OVERLAP 128      // chunks (2 KB per operand) from which a candidate gets the whole warp

struct RowArgs {
    const uint4 *bits;
    const i64 *bits_off;
    const uint2 *reg;
    const uint2 *span;
    const int4 *bbox;
    const u32 *area;
    const int *row_mask;
    const int *blk_grp;        // group of CTA b
    const int *blk_row0;       // first row of CTA b
    const int *grp_row_begin;
    const int *grp_row_count;
    const int *grp_col_begin;
    const int *grp_col_count;
    const i64 *grp_imat_off;
    int *imat;
    int *best_col;
    u32 *best_inter;
    double *best_score;
};

// popcount(A & B) over chunks lo + first, lo + first + stride, ... < hi; pointers are biased so that
// chunk c of a mask is at base[c].  Four 128-bit loads per operand in flight per lane.
__device__ __forceinline__ u32 strided_intersect(const uint4 *__restrict__ A, const uint4 *__restrict__ B,
                                                 u32 lo, u32 hi, u32 first, u32 stride)
{
    u32 acc = 0;
    for (u32 c = lo + first; c < hi; c += 4u * stride) {
        uint4 a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const u32 cc = c + stride * k;
            if (cc < hi) { a[k] = ld_v4_nc(A + cc); b[k] = ld_v4_nc(B + cc); }
            else { a[k] = make_uint4(0u, 0u, 0u, 0u); b[k] = a[k]; }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) acc += popc_and(a[k], b[k]);
    }
    return acc;
}

template <int MODE>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32, 4)
intersect_rows_kernel(const RowArgs p)
{
    __shared__ __align__(16) int4 s_bbox[COL_TILE];
    __shared__ __align__(16) uint2 s_span[COL_TILE + 2];
    __shared__ __align__(16) uint2 s_reg[COL_TILE + 2];
    __shared__ __align__(16) i64 s_off[COL_TILE + 2];
    __shared__ __align__(16) u32 s_area[COL_TILE + 4];
    __shared__ unsigned short s_cand[ROWS_PER_CTA][CAND_LIST];
    __shared__ __align__(8) u64 s_bar;

    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const int g = p.blk_grp[blockIdx.x];
    const int r = p.blk_row0[blockIdx.x] + (int)wid;
    const bool valid = r < p.grp_row_begin[g] + p.grp_row_count[g];
    const int cb = p.grp_col_begin[g];
    const int P = p.grp_col_count[g];
    const i64 imat_off = (p.imat && p.grp_imat_off) ? p.grp_imat_off[g] : -1;
    int *irow = (valid && imat_off >= 0) ? p.imat + imat_off + (i64)(r - p.grp_row_begin[g]) * P : nullptr;
    const u32 bar = smem_u32(&s_bar);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    int4 rb = make_int4(0, 0, -1, -1);
    uint2 rs = make_uint2(0u, 0u);
    u32 ra = 0;
    const uint4 *A = nullptr;
    if (valid) {
        const int rm = p.row_mask[r];
        rb = p.bbox[rm];
        rs = p.span[rm];
        ra = p.area[rm];
        A = p.bits + p.bits_off[rm] - p.reg[rm].x;
    }
    double best_s = 0.0;
    u32 best_i = 0;
    int best_c = MODE == AMPIS_MODE_IOU ? -1 : (P > 0 ? 0 : -1);
    unsigned short *list = s_cand[wid];
    u32 phase = 0;

    for (int t0 = 0; t0 < P; t0 += COL_TILE) {
        const int tn = min(COL_TILE, P - t0);
        const int c0m = cb + t0;                                   // first mask id of the tile
        const int a_skew = c0m & 3, o_skew = c0m & 1;              // element skew of the 16-byte aligned copies
        __syncthreads();                                           // previous tile consumed, barrier initialised
        // ---- stage the tile's metadata: TMA bulk copies for the 16-byte multiples, plain loads for the tails
        const int a_n = a_skew + tn, o_n = o_skew + tn;            // elements wanted from the aligned starts
        const int a_bulk = a_n & ~3, o_bulk = o_n & ~1;            // 4-byte items in fours, 8-byte items in pairs
        if (threadIdx.x == 0) {
            mbar_arrive_tx(bar, (u32)tn * 16u + (u32)a_bulk * 4u + 3u * (u32)o_bulk * 8u);
            bulk_g2s(smem_u32(s_bbox), p.bbox + c0m, (u32)tn * 16u, bar);
            if (a_bulk) bulk_g2s(smem_u32(s_area), p.area + (c0m - a_skew), (u32)a_bulk * 4u, bar);
            if (o_bulk) {
                bulk_g2s(smem_u32(s_off), p.bits_off + (c0m - o_skew), (u32)o_bulk * 8u, bar);
                bulk_g2s(smem_u32(s_span), p.span + (c0m - o_skew), (u32)o_bulk * 8u, bar);
                bulk_g2s(smem_u32(s_reg), p.reg + (c0m - o_skew), (u32)o_bulk * 8u, bar);
            }
        }
        if ((int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + (a_n - a_bulk))
            s_area[a_bulk + (int)threadIdx.x - 32] = p.area[c0m - a_skew + a_bulk + (int)threadIdx.x - 32];
        if ((int)threadIdx.x == 64 && o_n > o_bulk) {
            s_off[o_bulk] = p.bits_off[c0m - o_skew + o_bulk];
            s_span[o_bulk] = p.span[c0m - o_skew + o_bulk];
            s_reg[o_bulk] = p.reg[c0m - o_skew + o_bulk];
        }
        // dense row of this tile: zeros now, candidates patch their cells later
        if (irow) for (int k = (int)lane; k < tn; k += 32) irow[t0 + k] = 0;
        mbar_wait(bar, phase);
        phase ^= 1u;
        __syncthreads();                                           // tails written by plain stores
        if (!valid || ra == 0) continue;

        auto flush = [&](int n) {
            for (int j0 = 0; j0 < n; j0 += 4) {
                const int j = j0 + (int)(lane >> 3);
                const bool have = j < n;
                const int k = have ? (int)list[j] : 0;
                const uint2 cs = s_span[o_skew + k];
                const u32 lo = max(rs.x, cs.x), hi = min(rs.y, cs.y);
                const uint4 *B = p.bits + s_off[o_skew + k] - s_reg[o_skew + k].x;
                u32 inter = 0;
                if (__any_sync(0xffffffffu, have && hi - lo >= LONG_OVERLAP)) {
                    // long overlaps: the whole warp takes the four candidates one after another
                    for (int q = 0; q < 4 && j0 + q < n; q++) {
                        const int kq = (int)list[j0 + q];
                        const uint2 cq = s_span[o_skew + kq];
                        const u32 v = warp_sum(strided_intersect(A, p.bits + s_off[o_skew + kq] - s_reg[o_skew + kq].x,
                                                                 max(rs.x, cq.x), min(rs.y, cq.y), lane, 32));
                        if ((int)(lane >> 3) == q) inter = v;
                    }
                } else {
                    u32 v = have ? strided_intersect(A, B, lo, hi, lane & 7u, 8) : 0u;
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    inter = v;
                }
                if (have && (lane & 7u) == 0) {
                    const int c = t0 + k;
                    if (irow && inter) irow[c] = (int)inter;
                    if (MODE == AMPIS_MODE_IOU) {
                        // rleIou: u = a_r + a_c - i (the run walk's union); i == 0 => iou 0.0
                        const double s = inter ? (double)inter / (double)(ra + s_area[a_skew + k] - inter) : 0.0;
                        if (s > best_s || (s == best_s && s > 0.0 && (unsigned)c < (unsigned)best_c)) {
                            best_s = s; best_i = inter; best_c = c;
                        }
                    } else {
                        if (inter > best_i || (inter == best_i && inter > 0 && (unsigned)c < (unsigned)best_c)) {
                            best_i = inter; best_c = c;
                        }
                    }
                }
            }
        };

        int n = 0;
        for (int c0 = 0; c0 < tn; c0 += 32) {
            const int k = c0 + (int)lane;
            bool cand = false;
            if (k < tn) {
                const int4 b = s_bbox[k];
                const uint2 cs = s_span[o_skew + k];
                cand = max(rb.x, b.x) <= min(rb.z, b.z) && max(rb.y, b.y) <= min(rb.w, b.w) &&
                       max(rs.x, cs.x) < min(rs.y, cs.y);
            }
            const u32 bal = __ballot_sync(0xffffffffu, cand);
            if (cand) list[n + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)k;
            n += __popc(bal);
            if (n > CAND_LIST - 32) {                             // the next step could overflow the list
                __syncwarp();
                flush(n);
                n = 0;
                __syncwarp();
            }
        }
        __syncwarp();
        flush(n);
    }
    if (!valid) return;
    // warp arg-max: larger key wins, ties go to the smaller column index (np.argmax)
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
        if (MODE == AMPIS_MODE_SAT) best_s = (double)best_i / (double)ra;   // 0/0 = NaN like numpy
        p.best_col[r] = best_c;
        p.best_inter[r] = best_i;
        p.best_score[r] = best_s;
    }
}

extern "C" int ampis_intersect_rows(const void *d_bits, const int64_t *d_bits_off, const uint32_t *d_reg,
                                    const uint32_t *d_span, const int32_t *d_bbox, const uint32_t *d_area,
                                    const int32_t *d_row_mask, const int32_t *d_blk_grp,
                                    const int32_t *d_blk_row0, int32_t n_blocks,
                                    const int32_t *d_grp_row_begin, const int32_t *d_grp_row_count,
                                    const int32_t *d_grp_col_begin, const int32_t *d_grp_col_count,
                                    const int64_t *d_grp_imat_off, int32_t mode, int32_t *d_imat,
                                    int32_t *d_best_col, uint32_t *d_best_inter, double *d_best_score,
                                    void *stream)
{
    AMPIS_REQUIRE(n_blocks >= 0, "n_blocks < 0");
    AMPIS_REQUIRE(mode == AMPIS_MODE_IOU || mode == AMPIS_MODE_SAT, "bad mode");
    if (n_blocks == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits_off && d_reg && d_span && d_bbox && d_area && d_row_mask && d_blk_grp && d_blk_row0 &&
                      d_grp_row_begin && d_grp_row_count && d_grp_col_begin && d_grp_col_count && d_best_col &&
                      d_best_inter && d_best_score, "null pointer");
    AMPIS_REQUIRE((((uintptr_t)d_bbox | (uintptr_t)d_area | (uintptr_t)d_bits_off | (uintptr_t)d_span |
                    (uintptr_t)d_reg) & 15u) == 0,
                  "bbox / area / bits_off / span / reg must be 16-byte aligned (TMA bulk copies)");
    RowArgs a;
    a.bits = (const uint4 *)d_bits; a.bits_off = d_bits_off; a.reg = (const uint2 *)d_reg;
    a.span = (const uint2 *)d_span; a.bbox = (const int4 *)d_bbox; a.area = d_area;
    a.row_mask = d_row_mask; a.blk_grp = d_blk_grp; a.blk_row0 = d_blk_row0;
    a.grp_row_begin = d_grp_row_begin; a.grp_row_count = d_grp_row_count;
    a.grp_col_begin = d_grp_col_begin; a.grp_col_count = d_grp_col_count;
    a.grp_imat_off = d_grp_imat_off; a.imat = d_imat;
    a.best_col = d_best_col; a.best_inter = d_best_inter; a.best_score = d_best_score;
    if (mode == AMPIS_MODE_IOU)
        intersect_rows_kernel<AMPIS_MODE_IOU><<<n_blocks, ROWS_PER_CTA * 32, 0, as_stream(stream)>>>(a);
    else
        intersect_rows_kernel<AMPIS_MODE_SAT><<<n_blocks, ROWS_PER_CTA * 32, 0, as_stream(stream)>>>(a);
    AMPIS_CHECK_LAUNCH("intersect_rows_kernel");
    return AMPIS_OK;
}

extern "C" int ampis_rows_per_block(void) { return ROWS_PER_CTA; }

// float64 IoU matrix from dense intersections (analyze._piecewise_iou, analyze.py:54-112)
__global__ void __launch_bounds__(256)
iou_matrix_kernel(const int *__restrict__ imat, const u32 *__restrict__ ar, const u32 *__restrict__ ac, int G,
                  int P, double *__restrict__ out)
{
    const i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (i64)G * P) return;
    const int g = (int)(idx / P), c = (int)(idx - (i64)g * P);
    const u32 i = (u32)imat[idx];
    out[idx] = i ? (double)i / (double)(ar[g] + ac[c] - i) : 0.0;
}

extern "C" int ampis_iou_matrix_f64(const int32_t *d_imat, const uint32_t *d_area_rows,
                                    const uint32_t *d_area_cols, int32_t G, int32_t P, double *d_out,
                                    void *stream)
{
    AMPIS_REQUIRE(G >= 0 && P >= 0, "negative shape");
    if ((i64)G * P == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_imat && d_area_rows && d_area_cols && d_out, "null pointer");
    const i64 n = (i64)G * P;
    iou_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(d_imat, d_area_rows,
                                                                                 d_area_cols, G, P, d_out);
    AMPIS_CHECK_LAUNCH("iou_matrix_kernel");
    return AMPIS_OK;
}
