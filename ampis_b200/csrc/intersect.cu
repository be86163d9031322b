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
// A CTA owns up to 8 consecutive rows of ONE group (image); a warp owns one row.  The
// column masks' metadata (tight box, span, area, biased data pointer) is staged once per CTA
// into shared memory in tiles, so the 32-columns-at-a-time candidate scan of every warp runs
// out of shared memory instead of paying a global round trip per step.  Candidates (boxes
// overlap AND spans overlap; pruning cannot change a result: disjoint boxes => intersection 0
// => score 0) are then intersected one after another by the whole warp: 128-bit loads of both
// packed masks over the overlap of their spans, AND + popc, warp reduction.
#include "common.cuh"

#define ROWS_PER_CTA 8
#define COL_TILE 512

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

// popcount(A & B) over chunks [lo,hi); pointers are biased so that chunk c of a mask is at
// base[c].  Whole-warp cooperative; up to 4 x 128-bit loads per operand in flight per lane,
// predicated so that ranges up to 128 chunks (2 KB per operand) take ONE memory round trip.
__device__ __forceinline__ u32 warp_intersect(const uint4 *__restrict__ A, const uint4 *__restrict__ B,
                                              u32 lo, u32 hi, u32 lane)
{
    u32 acc = 0;
    for (u32 c = lo + lane; c < hi; c += 128) {
        uint4 a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const u32 cc = c + 32u * k;
            if (cc < hi) { a[k] = ld_v4_nc(A + cc); b[k] = ld_v4_nc(B + cc); }
            else { a[k] = make_uint4(0u, 0u, 0u, 0u); b[k] = a[k]; }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) acc += popc_and(a[k], b[k]);
    }
    return warp_sum(acc);
}

template <int MODE>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32, 6)
intersect_rows_kernel(const RowArgs p)
{
    __shared__ int4 s_bbox[COL_TILE];
    __shared__ uint2 s_span[COL_TILE];
    __shared__ u32 s_area[COL_TILE];
    __shared__ const uint4 *s_base[COL_TILE];

    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const int g = p.blk_grp[blockIdx.x];
    const int r = p.blk_row0[blockIdx.x] + (int)wid;
    const bool valid = r < p.grp_row_begin[g] + p.grp_row_count[g];
    const int cb = p.grp_col_begin[g];
    const int P = p.grp_col_count[g];
    const i64 imat_off = (p.imat && p.grp_imat_off) ? p.grp_imat_off[g] : -1;
    int *irow = (valid && imat_off >= 0) ? p.imat + imat_off + (i64)(r - p.grp_row_begin[g]) * P : nullptr;

    int4 rb = make_int4(0, 0, -1, -1);
    uint2 rs = make_uint2(0u, 0u);
    u32 ra = 0;
    const uint4 *A = nullptr;
    if (valid) {
        const int rm = p.row_mask[r];
        rb = p.bbox[rm];
        ra = p.area[rm];
        rs = p.span[rm];
        A = p.bits + p.bits_off[rm] - p.reg[rm].x;
    }

    // lane-local running best over the columns this lane owns (increasing index => first max)
    double best_s = 0.0;
    u32 best_i = 0;
    int best_c = MODE == AMPIS_MODE_IOU ? -1 : (P > 0 ? 0 : -1);

    for (int t0 = 0; t0 < P; t0 += COL_TILE) {
        const int tn = min(COL_TILE, P - t0);
        __syncthreads();   // previous tile fully consumed
        for (int k = threadIdx.x; k < tn; k += blockDim.x) {
            const int cm = cb + t0 + k;
            s_bbox[k] = p.bbox[cm];
            s_area[k] = p.area[cm];
            s_span[k] = p.span[cm];
            s_base[k] = p.bits + p.bits_off[cm] - p.reg[cm].x;
        }
        __syncthreads();
        if (!valid) continue;
        for (int c0 = 0; c0 < tn; c0 += 32) {
            const int k = c0 + (int)lane;
            bool cand = false;
            uint2 cs = make_uint2(0u, 0u);
            u32 ca = 0;
            if (k < tn && ra > 0) {
                const int4 b = s_bbox[k];
                ca = s_area[k];
                cs = s_span[k];
                cand = max(rb.x, b.x) <= min(rb.z, b.z) && max(rb.y, b.y) <= min(rb.w, b.w) &&
                       max(rs.x, cs.x) < min(rs.y, cs.y);
            }
            u32 inter = 0;
            u32 todo = __ballot_sync(0xffffffffu, cand);
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const uint2 ss = s_span[c0 + src];
                const u32 v = warp_intersect(A, s_base[c0 + src], max(rs.x, ss.x), min(rs.y, ss.y), lane);
                if ((int)lane == src) inter = v;
            }
            if (k < tn) {
                const int c = t0 + k;
                if (irow) irow[c] = (int)inter;
                if (MODE == AMPIS_MODE_IOU) {
                    // rleIou: u = a_r + a_c - i (the run walk's union); i == 0 => iou 0.0
                    const double s = inter ? (double)inter / (double)(ra + ca - inter) : 0.0;
                    if (s > best_s) { best_s = s; best_i = inter; best_c = c; }
                } else {
                    if (inter > best_i) { best_i = inter; best_c = c; }
                }
            }
        }
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
