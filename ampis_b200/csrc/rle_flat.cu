// RLE -> bounding-box windows (AMPIS_LAYOUT_CROP), third generation: a warp decodes SEVERAL masks at once with
// its lanes spread over the concatenated run lists ("flat"), instead of a group of 8 / 16 / 32 lanes per mask.
//
// Same role and same outputs as rle_measure_paint_crop_kernel (rle_paint.cu): rleArea (structures.py:568,571;
// analyze.py:320-321; powder.py:264), the tight box extract_boxes reads off the decoded mask (data_utils.py:229-239)
// and rleDecode (structures.py:752,761) restricted to that box -- what every RLE.iou / RLE.merge call of
// analyze.py:158,315 and powder.py:82 decodes again and again in the reference.
//
// Why: ncu on the group-per-mask kernel (profiles/kernels_r01p.md) showed it bound by instruction issue at ~740
// warp instructions per mask.  Most of them were per-MASK fixed costs executed by a whole group (five shuffle
// reductions, the 60-instruction record store, arena reservation, tile zero-fill / copy-out loops with 1-2 trips)
// and passes in which half of the lanes had no run left (75 runs on 2 x 64 slots).  Here
//   * the unit of lane work is a PAIR (0-run, 1-run); the K masks a warp owns are laid end to end, so every pass of
//     32 pairs is full except the last one of the warp;
//   * run end positions come from ONE segmented warp scan per pass (32-bit saturating arithmetic: frames of
//     2^31 pixels or more go to the fallback kernel);
//   * per-mask statistics are reduced with REDUX over the lanes of each mask present in the pass (1-2 masks);
//   * the per-mask records (area, box, span, region, status, arena offset) are formed by ONE lane per mask, all
//     masks of the warp in parallel, and the warp reserves arena space for all of them with a single atomicAdd;
//   * the windows of all K masks are assembled in one shared-memory tile that is contiguous in the arena, so the
//     copy-out is a straight run of 128-bit stores.
// Masks that do not fit the per-warp budgets (more than FL_PAIR_CAP pairs, window larger than the tile, frame of
// 2^31 pixels or more) are appended to a list that rle_measure_paint_list_kernel (a warp per mask, any size;
// rle_paint.cu) works off in a second launch.
#include "common.cuh"
#include "rle_measure.cuh"

#define FL_WARPS 8
#define FL_KMAX 16              // masks per warp and round
#define FL_PAIR_CAP 256         // (0-run, 1-run) pairs per warp and round
#define FL_TILE_WORDS 512       // window words per warp and round
#define FL_SAT 0x7fffffffu      // positions saturate here; frames must be smaller
#define FL_NONE 0xffffffffu

struct __align__(16) FlatWarp {
    uint2 pair[FL_PAIR_CAP];        // (start, length) of the 1-run of every pair; length 0 = nothing to paint
    u32 tile[FL_TILE_WORDS];
    int4 bb[FL_KMAX];
    i64 off[FL_KMAX];
    int len[FL_KMAX];
    u32 H[FL_KMAX], rcp[FL_KMAX], HW[FL_KMAX];
    u32 pb[FL_KMAX + 1];            // first pair of every mask of the round
    u32 area[FL_KMAX], first[FL_KMAX], last[FL_KMAX], ymin[FL_KMAX], ymax[FL_KMAX], total[FL_KMAX];
    u32 woff[FL_KMAX];              // first tile word of the mask's window, FL_NONE = not painted here
    uint8_t mid[FL_PAIR_CAP];       // mask (index inside the round) of every pair
};

__device__ __forceinline__ u32 sat_add(u32 a, u32 b) { return min(a + b, FL_SAT); }

__device__ __forceinline__ u32 warp_incl_scan(u32 v, u32 lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, v, d);
        if ((int)lane >= d) v += t;
    }
    return v;
}

__device__ __forceinline__ u32 div_by(u32 s, u32 d, u32 rcp)
{
    u32 q = __umulhi(s, rcp);
    if (s - q * d >= d) q++;
    return q;
}

__global__ void __launch_bounds__(FL_WARPS * 32, 5)
rle_flat_crop_kernel(const u32 *__restrict__ cnt, const i64 *__restrict__ cnt_off, const int *__restrict__ cnt_len,
                     const u32 *__restrict__ hh, const u32 *__restrict__ ww, int n, int K,
                     u32 *__restrict__ area, int *__restrict__ bbox, u32 *__restrict__ span, u32 *__restrict__ reg,
                     i64 *__restrict__ bits_off, int *__restrict__ status, uint4 *__restrict__ bits, i64 capacity,
                     unsigned long long *__restrict__ cursor, int *__restrict__ big)
{
    __shared__ FlatWarp s_warp[FL_WARPS];
    FlatWarp &S = s_warp[threadIdx.x >> 5];
    const u32 lane = lane_id();
    const i64 gw = (i64)blockIdx.x * FL_WARPS + (threadIdx.x >> 5);
    const i64 i0 = gw * K;
    if (i0 >= n) return;
    const int nm = (int)min((i64)K, (i64)n - i0);

    // ---- the warp's masks: one lane each
    u32 my_pairs = 0;
    bool my_big = false;
    if ((int)lane < nm) {
        const i64 i = i0 + lane;
        const int len = cnt_len[i];
        const u32 H = hh[i];
        const u64 HW = (u64)H * ww[i];
        S.len[lane] = len;
        S.off[lane] = cnt_off[i];
        S.H[lane] = H;
        S.rcp[lane] = H ? 0xffffffffu / H : 0u;
        S.HW[lane] = (u32)HW;
        my_pairs = (u32)(len + 1) >> 1;
        my_big = len < 0 || H == 0 || HW >= (u64)FL_SAT || my_pairs > FL_PAIR_CAP;
    }
    __syncwarp();

    int a = 0;                                       // first mask of the round (index inside the warp's masks)
    while (a < nm) {
        // ---- round = longest prefix of the remaining masks whose pairs fit (a single small mask always does)
        const bool mine = (int)lane >= a && (int)lane < nm;
        const u32 v = (mine && !my_big) ? my_pairs : 0u;
        const u32 incl = warp_incl_scan(v, lane);
        const bool in_round = mine && incl <= FL_PAIR_CAP;
        const int nb = __popc(__ballot_sync(0xffffffffu, in_round));
        const u32 T = __shfl_sync(0xffffffffu, incl, a + nb - 1);
        const int j_me = (int)lane - a;               // my mask's index inside the round
        if (in_round) {
            S.pb[j_me] = incl - v;
            if (j_me == nb - 1) S.pb[nb] = T;
            S.area[j_me] = 0u; S.first[j_me] = 0xffffffffu; S.last[j_me] = 0u;
            S.ymin[j_me] = 0xffffffffu; S.ymax[j_me] = 0u; S.total[j_me] = 0u;
            S.woff[j_me] = FL_NONE;
            if (my_big) big[1 + atomicAdd(&big[0], 1)] = (int)(i0 + lane);
        }
        __syncwarp();

        // ---- pass over the pairs: run end positions by a segmented scan, statistics of the 1-runs
        u32 carry = 0;
        int j = 0;                                   // my mask: the last one that starts at or before my pair; only grows
        for (u32 f0 = 0; f0 < T; f0 += 32) {
            const u32 f = f0 + lane;
            const bool ok = f < T;
            u32 p = 0, z = 0, o = 0, H = 1, rcp = 0, HW = 0;
            int len = 0;
            if (ok) {
                while (j + 1 < nb && f >= S.pb[j + 1]) j++;
                S.mid[f] = (uint8_t)j;
                p = f - S.pb[j];
                len = S.len[a + j];
                const u32 *c = cnt + S.off[a + j] + 2 * p;
                z = min(__ldg(c), FL_SAT);
                if ((int)(2 * p + 1) < len) o = min(__ldg(c + 1), FL_SAT);
                H = S.H[a + j]; rcp = S.rcp[a + j]; HW = S.HW[a + j];
            }
            u32 end = sat_add(z, o);
            const u32 reach = ok ? min(lane, p) : 0u;            // lanes below me that belong to my mask
#pragma unroll
            for (u32 d = 1; d < 32; d <<= 1) {
                const u32 t = __shfl_up_sync(0xffffffffu, end, d);
                if (d <= reach) end = sat_add(end, t);
            }
            if (ok && p > lane) end = sat_add(end, carry);        // my mask began in an earlier pass
            carry = __shfl_sync(0xffffffffu, end, 31);
            const bool good = ok && o > 0u && end <= HW;           // HW < FL_SAT: nothing saturated up to here
            const u32 start = end - o;
            if (ok) {
                S.pair[f] = good ? make_uint2(start, o) : make_uint2(0u, 0u);
                if (p == ((u32)(len + 1) >> 1) - 1u) S.total[j] = end;
            }
            u32 s_a = 0, s_f = 0xffffffffu, s_l = 0, s_y0 = 0xffffffffu, s_y1 = 0;
            if (good) {
                const u32 xs = div_by(start, H, rcp), xe = div_by(end - 1u, H, rcp);
                s_a = o; s_f = start; s_l = end;
                if (xs != xe) { s_y0 = 0u; s_y1 = H - 1u; }
                else { s_y0 = start - xs * H; s_y1 = end - 1u - xe * H; }
            }
            // reduce over the lanes of every mask present in this pass (masks are contiguous lane ranges)
            u32 todo = __ballot_sync(0xffffffffu, ok);
            while (todo) {
                const int src = __ffs(todo) - 1;
                const int j0 = __shfl_sync(0xffffffffu, j, src);
                const u32 lm = __ballot_sync(0xffffffffu, ok && j == j0);
                todo &= ~lm;
                if (ok && j == j0) {
                    const u32 r_a = __reduce_add_sync(lm, s_a);
                    const u32 r_f = __reduce_min_sync(lm, s_f);
                    const u32 r_l = __reduce_max_sync(lm, s_l);
                    const u32 r_y0 = __reduce_min_sync(lm, s_y0);
                    const u32 r_y1 = __reduce_max_sync(lm, s_y1);
                    if ((int)lane == src) {
                        S.area[j0] += r_a;
                        S.first[j0] = min(S.first[j0], r_f);
                        S.last[j0] = max(S.last[j0], r_l);
                        S.ymin[j0] = min(S.ymin[j0], r_y0);
                        S.ymax[j0] = max(S.ymax[j0], r_y1);
                    }
                }
            }
        }
        __syncwarp();

        // ---- records: one lane per mask; window sizes -> tile offsets and ONE arena reservation per warp
        u32 chunks = 0;
        int4 bb = make_int4(0, 0, -1, -1);
        const bool rec = in_round && !my_big;
        if (rec) {
            const u32 H = S.H[lane], HW = S.HW[lane];
            const u32 ar = S.area[j_me], fi = S.first[j_me], la = S.last[j_me];
            const u32 nchunks = (HW + AMPIS_CHUNK_BITS - 1u) / AMPIS_CHUNK_BITS;
            uint2 sp = make_uint2(0u, 0u);
            if (ar > 0u) {
                const u32 rcp = S.rcp[lane];
                sp = make_uint2(fi / AMPIS_CHUNK_BITS, min((la + AMPIS_CHUNK_BITS - 1u) / AMPIS_CHUNK_BITS, nchunks));
                bb = make_int4((int)div_by(fi, H, rcp), (int)S.ymin[j_me], (int)div_by(la - 1u, H, rcp), (int)S.ymax[j_me]);
            }
            chunks = (crop_words(bb) + 3u) / 4u;
            const i64 i = i0 + lane;
            area[i] = ar;
            reinterpret_cast<int4 *>(bbox)[i] = bb;
            reinterpret_cast<uint2 *>(span)[i] = sp;
            reinterpret_cast<uint2 *>(reg)[i] = make_uint2(0u, chunks);
            status[i] = S.total[j_me] == HW ? 0 : AMPIS_ST_BAD_TOTAL;
            S.bb[j_me] = bb;
        }
        const u32 cincl = warp_incl_scan(chunks, lane);
        __syncwarp();

        // ---- windows: the round's masks are painted in groups that fit the tile (usually all of them at once).
        // Per group: ONE arena reservation, the tile is zeroed, every lane sets the bits of its 1-runs, and the tile
        // -- contiguous in the arena -- is copied out with 128-bit stores.
        int lo = a;                                       // first mask (lane) of the round not yet painted
        const int hi = a + nb;
        while (lo < hi) {
            const u32 before = lo > 0 ? __shfl_sync(0xffffffffu, cincl, lo - 1) : 0u;
            const u32 rel = cincl - before;                // chunks of the masks lo..me
            const bool here = (int)lane >= lo && (int)lane < hi && rel * 4u <= FL_TILE_WORDS;       // a prefix of lo..hi-1
            const u32 here_m = __ballot_sync(0xffffffffu, here);
            if (!((here_m >> lo) & 1u)) {                 // this window alone is larger than the tile: fallback kernel
                if ((int)lane == lo) big[1 + atomicAdd(&big[0], 1)] = (int)(i0 + lane);
                lo++;
                continue;
            }
            const int last = 31 - __clz(here_m);
            const u32 fit = __shfl_sync(0xffffffffu, rel, last);
            unsigned long long base = 0;
            if (lane == 0 && fit) base = atomicAdd(cursor, (unsigned long long)fit);
            base = __shfl_sync(0xffffffffu, base, 0);
            const bool room = (i64)(base + fit) <= capacity;
            if (here && rec) {
                const i64 i = i0 + lane;
                if (room) {
                    bits_off[i] = (i64)base + (rel - chunks);
                    S.woff[j_me] = (rel - chunks) * 4u;
                } else {
                    // arena exhausted: the caller sees *cursor > capacity and retries; until then the mask is empty
                    // for every later kernel (the crop rows kernels size their reads from the box)
                    bits_off[i] = 0;
                    reinterpret_cast<int4 *>(bbox)[i] = make_int4(0, 0, -1, -1);
                    reinterpret_cast<uint2 *>(span)[i] = make_uint2(0u, 0u);
                    reinterpret_cast<uint2 *>(reg)[i] = make_uint2(0u, 0u);
                }
            }
            __syncwarp();
            if (fit && room) {
                uint4 *tile4 = reinterpret_cast<uint4 *>(S.tile);
                for (u32 k = lane; k < fit; k += 32) tile4[k] = make_uint4(0u, 0u, 0u, 0u);
                __syncwarp();
                const u32 f_end = S.pb[last - a + 1];
                for (u32 f = S.pb[lo - a] + lane; f < f_end; f += 32) {
                    const uint2 pr = S.pair[f];
                    if (pr.y == 0u) continue;
                    const int j = S.mid[f];
                    const u32 wo = S.woff[j];
                    const int4 b = S.bb[j];
                    const u32 H = S.H[a + j];
                    const u32 wy0 = (u32)b.y >> 5, nwy = ((u32)b.w >> 5) - wy0 + 1u;
                    u32 s = pr.x;
                    const u32 e = pr.x + pr.y;
                    u32 x = div_by(s, H, S.rcp[a + j]);
                    u32 cs = x * H;
                    while (s < e) {
                        // rows [ys,ye) of column x, clipped to the box (a well-formed mask never needs the clip)
                        const u32 ys = max(s - cs, (u32)b.y), ye = min(min(e, cs + H) - cs, (u32)b.w + 1u);
                        if (ye > ys && x >= (u32)b.x && x <= (u32)b.z) {
                            u32 *col = S.tile + wo + (x - (u32)b.x) * nwy - wy0;
                            const u32 w0 = ys >> 5, w1 = (ye - 1u) >> 5;
                            const u32 m0 = 0xffffffffu << (ys & 31u), m1 = 0xffffffffu >> (31u - ((ye - 1u) & 31u));
                            if (w0 == w1) {
                                atomicOr(&col[w0], m0 & m1);
                            } else {
                                atomicOr(&col[w0], m0);
                                for (u32 w = w0 + 1u; w < w1; w++) col[w] = 0xffffffffu;
                                atomicOr(&col[w1], m1);
                            }
                        }
                        cs += H; s = cs; x++;
                    }
                }
                __syncwarp();
                uint4 *out = bits + base;
                for (u32 k = lane; k < fit; k += 32) out[k] = tile4[k];
            }
            __syncwarp();
            lo = last + 1;
        }
        a += nb;
    }
}

// masks per warp for a typical number of runs per mask: fill ~4/5 of the pair budget (a warp whose masks exceed
// it simply works in two rounds)
static int flat_masks_per_warp(int runs_hint)
{
    if (runs_hint <= 0) return 4;
    const int pairs = runs_hint / 2 + 1;
    int k = (FL_PAIR_CAP * 4 / 5) / pairs;
    if (k < 1) k = 1;
    if (k > FL_KMAX) k = FL_KMAX;
    return k;
}

int ampis_launch_measure_paint_list(const uint32_t *d_cnt, const int64_t *d_cnt_off, const int32_t *d_cnt_len,
                                    const uint32_t *d_h, const uint32_t *d_w, const int32_t *d_list, int32_t max_n,
                                    uint32_t *d_cum, uint32_t *d_area, int32_t *d_bbox, uint32_t *d_span,
                                    uint32_t *d_reg, int64_t *d_bits_off, int32_t *d_status, void *d_bits,
                                    int64_t bits_capacity, uint64_t *d_cursor, cudaStream_t st);

extern "C" int ampis_rle_measure_paint_flat(const uint32_t *d_cnt, const int64_t *d_cnt_off, const int32_t *d_cnt_len,
                                            const uint32_t *d_h, const uint32_t *d_w, int32_t n, uint32_t *d_cum,
                                            uint32_t *d_area, int32_t *d_bbox, uint32_t *d_span, uint32_t *d_reg,
                                            int64_t *d_bits_off, int32_t *d_status, void *d_bits,
                                            int64_t bits_capacity, uint64_t *d_cursor, int32_t *d_list,
                                            int32_t runs_hint, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_cnt && d_cnt_off && d_cnt_len && d_h && d_w && d_cum && d_area && d_bbox && d_span && d_reg &&
                      d_bits_off && d_status && d_bits && d_cursor && d_list, "null pointer");
    AMPIS_REQUIRE(((uintptr_t)d_bits & 15u) == 0, "bits arena must be 16-byte aligned");
    cudaStream_t st = as_stream(stream);
    cudaError_t e = cudaMemsetAsync(d_cursor, 0, sizeof(uint64_t), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_list, 0, sizeof(int32_t), st);
    if (e != cudaSuccess) { ampis_set_error("cursor memset: %s", cudaGetErrorString(e)); return AMPIS_ECUDA; }
    const int K = flat_masks_per_warp(runs_hint);
    const int64_t warps = ((int64_t)n + K - 1) / K;
    const unsigned grid = (unsigned)((warps + FL_WARPS - 1) / FL_WARPS);
    rle_flat_crop_kernel<<<grid, FL_WARPS * 32, 0, st>>>(
        d_cnt, d_cnt_off, d_cnt_len, d_h, d_w, n, K, d_area, d_bbox, d_span, d_reg, d_bits_off, d_status,
        (uint4 *)d_bits, bits_capacity, (unsigned long long *)d_cursor, d_list);
    AMPIS_CHECK_LAUNCH("rle_flat_crop_kernel");
    return ampis_launch_measure_paint_list(d_cnt, d_cnt_off, d_cnt_len, d_h, d_w, d_list, n, d_cum, d_area, d_bbox,
                                           d_span, d_reg, d_bits_off, d_status, d_bits, bits_capacity, d_cursor, st);
}
