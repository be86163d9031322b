// Intersection rows over AMPIS_LAYOUT_CROP tables, second generation (same outputs as
// intersect_rows_kernel: analyze.py:149-164 / powder.py:80-86 semantics, see intersect.cu).
//
// What bounded the first version (ncu source-level stall sampling, round 1): 61 % of the stall samples were loads
// -- every candidate pair cost one exposed global round trip, one after another, and every CTA
// re-staged the column metadata through registers.  Here
//   * the column metadata of the image (tight boxes, areas, arena offsets: three contiguous arrays)
//     is staged into shared memory by TMA bulk copies (cp.async.bulk + mbarrier complete_tx), one
//     elected thread, no per-element address arithmetic;
//   * a warp first scans ALL columns of the tile against its row (shared memory only) and collects
//     the candidates in a list, then intersects them four at a time, eight lanes per candidate,
//     so the loads of four candidates (and of consecutive words of each) are in flight together;
//     candidates with large overlaps take the whole warp;
//   * the dense row of the intersection matrix is zero-filled with coalesced stores during the scan,
//     candidates patch their cell afterwards.
#include "common.cuh"
#include "async.cuh"
#include "crop_common.cuh"

#define RC_ROWS AMPIS_ROWS_PER_CTA
#define RC_TILE 512
#define RC_LIST 64
#define RC_BIG 256          // overlap words from which a candidate gets the whole warp

struct CropRowArgs {
    const u32 *words;
    const i64 *bits_off;
    const int4 *bbox;
    const u32 *area;
    const int *row_mask;
    const int *blk_grp, *blk_row0;
    const int *grp_row_begin, *grp_row_count, *grp_col_begin, *grp_col_count;
    const i64 *grp_imat_off;
    int *imat;
    int *best_col;
    u32 *best_inter;
    double *best_score;
};

template <int MODE>
__global__ void __launch_bounds__(RC_ROWS * 32, 4)
intersect_rows_crop_kernel(const CropRowArgs p)
{
    __shared__ __align__(16) int4 s_bbox[RC_TILE];
    __shared__ __align__(16) u32 s_area[RC_TILE + 4];
    __shared__ __align__(16) i64 s_off[RC_TILE + 2];
    __shared__ unsigned short s_cand[RC_ROWS][RC_LIST];
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
    u32 ra = 0;
    const u32 *A = nullptr;
    if (valid) {
        const int rm = p.row_mask[r];
        rb = p.bbox[rm];
        ra = p.area[rm];
        A = p.words + p.bits_off[rm] * 4;
    }
    double best_s = 0.0;
    u32 best_i = 0;
    int best_c = MODE == AMPIS_MODE_IOU ? -1 : (P > 0 ? 0 : -1);
    unsigned short *list = s_cand[wid];
    u32 phase = 0;

    for (int t0 = 0; t0 < P; t0 += RC_TILE) {
        const int tn = min(RC_TILE, P - t0);
        const int c0m = cb + t0;                                   // first mask id of the tile
        const int a_skew = c0m & 3, o_skew = c0m & 1;              // element skew of the 16-byte aligned copies
        __syncthreads();                                           // previous tile consumed, barrier initialised
        // ---- stage the tile's metadata: TMA bulk copies for the 16-byte multiples, plain loads for the tails
        const int a_n = a_skew + tn, o_n = o_skew + tn;            // elements wanted from the aligned starts
        const int a_bulk = a_n & ~3, o_bulk = o_n & ~1;
        if (threadIdx.x == 0) {
            const u32 bytes = (u32)tn * 16u + (u32)a_bulk * 4u + (u32)o_bulk * 8u;
            mbar_arrive_tx(bar, bytes);
            bulk_g2s(smem_u32(s_bbox), p.bbox + c0m, (u32)tn * 16u, bar);
            if (a_bulk) bulk_g2s(smem_u32(s_area), p.area + (c0m - a_skew), (u32)a_bulk * 4u, bar);
            if (o_bulk) bulk_g2s(smem_u32(s_off), p.bits_off + (c0m - o_skew), (u32)o_bulk * 8u, bar);
        }
        if ((int)threadIdx.x >= 32 && (int)threadIdx.x < 32 + (a_n - a_bulk))
            s_area[a_bulk + (int)threadIdx.x - 32] = p.area[c0m - a_skew + a_bulk + (int)threadIdx.x - 32];
        if ((int)threadIdx.x == 64 && o_n > o_bulk) s_off[o_bulk] = p.bits_off[c0m - o_skew + o_bulk];
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
                Overlap o;
                o.total = 0;
                if (have) o = overlap_of(A, rb, p.words + s_off[o_skew + k] * 4, s_bbox[k]);
                u32 inter = 0;
                if (__any_sync(0xffffffffu, have && o.total >= RC_BIG)) {
                    // large overlaps: the whole warp takes the four candidates one after another
                    for (int q = 0; q < 4 && j0 + q < n; q++) {
                        const int kq = (int)list[j0 + q];
                        const Overlap oq = overlap_of(A, rb, p.words + s_off[o_skew + kq] * 4, s_bbox[kq]);
                        const u32 v = warp_sum(overlap_popc(oq, lane, 32));
                        if ((int)(lane >> 3) == q) inter = v;
                    }
                } else {
                    u32 v = have ? overlap_popc(o, lane & 7u, 8) : 0u;
                    v += __shfl_xor_sync(0xffffffffu, v, 4);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    inter = v;
                }
                if (have && (lane & 7u) == 0) {
                    const int c = t0 + k;
                    if (irow && inter) irow[c] = (int)inter;
                    if (MODE == AMPIS_MODE_IOU) {
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
                cand = b.x <= rb.z && b.z >= rb.x && b.y <= rb.w && b.w >= rb.y;      // boxes are non-empty here
            }
            const u32 bal = __ballot_sync(0xffffffffu, cand);
            if (cand) list[n + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)k;
            n += __popc(bal);
            if (n > RC_LIST - 32) {                               // the next step could overflow the list
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

extern "C" int ampis_intersect_rows_crop(const void *d_bits, const int64_t *d_bits_off, const int32_t *d_bbox,
                                         const uint32_t *d_area, const int32_t *d_row_mask,
                                         const int32_t *d_blk_grp, const int32_t *d_blk_row0, int32_t n_blocks,
                                         const int32_t *d_grp_row_begin, const int32_t *d_grp_row_count,
                                         const int32_t *d_grp_col_begin, const int32_t *d_grp_col_count,
                                         const int64_t *d_grp_imat_off, int32_t mode, int32_t *d_imat,
                                         int32_t *d_best_col, uint32_t *d_best_inter, double *d_best_score,
                                         void *stream)
{
    AMPIS_REQUIRE(n_blocks >= 0, "n_blocks < 0");
    AMPIS_REQUIRE(mode == AMPIS_MODE_IOU || mode == AMPIS_MODE_SAT, "bad mode");
    if (n_blocks == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits_off && d_bbox && d_area && d_row_mask && d_blk_grp && d_blk_row0 && d_grp_row_begin &&
                      d_grp_row_count && d_grp_col_begin && d_grp_col_count && d_best_col && d_best_inter &&
                      d_best_score, "null pointer");
    AMPIS_REQUIRE((((uintptr_t)d_bbox | (uintptr_t)d_area | (uintptr_t)d_bits_off) & 15u) == 0,
                  "bbox / area / bits_off must be 16-byte aligned (TMA bulk copies)");
    CropRowArgs a;
    a.words = (const u32 *)d_bits; a.bits_off = d_bits_off; a.bbox = (const int4 *)d_bbox; a.area = d_area;
    a.row_mask = d_row_mask; a.blk_grp = d_blk_grp; a.blk_row0 = d_blk_row0;
    a.grp_row_begin = d_grp_row_begin; a.grp_row_count = d_grp_row_count;
    a.grp_col_begin = d_grp_col_begin; a.grp_col_count = d_grp_col_count;
    a.grp_imat_off = d_grp_imat_off; a.imat = d_imat;
    a.best_col = d_best_col; a.best_inter = d_best_inter; a.best_score = d_best_score;
    if (mode == AMPIS_MODE_IOU)
        intersect_rows_crop_kernel<AMPIS_MODE_IOU><<<n_blocks, RC_ROWS * 32, 0, as_stream(stream)>>>(a);
    else
        intersect_rows_crop_kernel<AMPIS_MODE_SAT><<<n_blocks, RC_ROWS * 32, 0, as_stream(stream)>>>(a);
    AMPIS_CHECK_LAUNCH("intersect_rows_crop_kernel");
    return AMPIS_OK;
}
