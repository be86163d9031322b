// RLE -> bit-packed masks ("paint"), and conversions between packed masks and
// bool[n][h][w] arrays.
//
// Layout: a mask is the h*w-bit vector of its pixels in COCO's column-major order
// (bit k = pixel x*h+y), cut into 128-bit chunks (uint4).  Only the chunk region
// [reg_lo,reg_hi) of a mask is stored (SPAN layout: first..last 1-pixel; FULL layout:
// the whole frame).  Because every mask of an image uses the same linear index, the
// intersection of two masks is AND+popc over the overlap of their regions with no shifts
// and no transposes; only masks_to_bitmask_array needs the row-major view (unpack below).
#include "common.cuh"
#include "rle_measure.cuh"

// bits [lo,hi) of a 32-bit word, 0 <= lo <= hi <= 32
__device__ __forceinline__ u32 bit_range(u32 lo, u32 hi)
{
    const u32 upto_hi = hi >= 32 ? 0xffffffffu : ((1u << hi) - 1u);
    const u32 upto_lo = lo >= 32 ? 0xffffffffu : ((1u << lo) - 1u);
    return upto_hi & ~upto_lo;
}

// The 128 pixels [128c, 128c+128) of a mask whose run END positions are C[0..m).
// LDG: C is read-only global memory written by an earlier kernel (non-coherent loads allowed);
// otherwise C is shared memory or global memory written earlier in this kernel.
template <bool LDG>
__device__ __forceinline__ uint4 paint_chunk(const u32 *C, int m, u32 c)
{
    const u64 b0 = (u64)c * AMPIS_CHUNK_BITS, b1 = b0 + AMPIS_CHUNK_BITS;
    // first run r whose end lies beyond b0 -- the run that owns pixel b0
    int lo = 0, hi = m;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((u64)(LDG ? __ldg(C + mid) : C[mid]) > b0) hi = mid; else lo = mid + 1;
    }
    int r = lo;
    u32 w[4] = {0u, 0u, 0u, 0u};
    u64 pos = b0;
    while (r < m && pos < b1) {
        const u64 e = min((u64)(LDG ? __ldg(C + r) : C[r]), b1);
        if (r & 1) {
            const u32 s = (u32)(pos - b0), t = (u32)(e - b0);   // [s,t) within the chunk
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const u32 ws = 32u * k;
                const u32 a = s > ws ? s - ws : 0u;
                const u32 b = t > ws ? min(t - ws, 32u) : 0u;
                if (b > a) w[k] |= bit_range(a, b);
            }
        }
        pos = e;
        r++;
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// One CTA per mask (grid-stride): threads stride over the mask's region so that a warp
// stores 512 contiguous bytes per instruction.  Chunks outside the span are zeros and cost
// no search; inside the span each chunk is located by a binary search over the run ends
// (L1-resident: ~0.5 KB per mask).
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
rle_paint_kernel(const u32 *__restrict__ cum, const i64 *__restrict__ cnt_off, const int *__restrict__ cnt_len,
                 const uint2 *__restrict__ span, const uint2 *__restrict__ reg,
                 const i64 *__restrict__ bits_off, int n, uint4 *__restrict__ bits, i64 capacity)
{
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint2 rg = reg[i];
        const i64 off = bits_off[i];
        if (off + (i64)(rg.y - rg.x) > capacity) continue;
        const uint2 sp = span[i];
        const int m = cnt_len[i];
        const u32 *C = cum + cnt_off[i];
        uint4 *out = bits + off - rg.x;
        for (u32 c = rg.x + threadIdx.x; c < rg.y; c += THREADS) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (c >= sp.x && c < sp.y) v = paint_chunk<true>(C, m, c);
            st_v4_stream(out + c, v);
        }
    }
}

extern "C" int ampis_rle_decode_packed(const uint32_t *d_cum, const int64_t *d_cnt_off,
                                       const int32_t *d_cnt_len, const uint32_t *d_span, const uint32_t *d_reg,
                                       const int64_t *d_bits_off, int32_t n, void *d_bits, int64_t bits_capacity,
                                       void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_cum && d_cnt_off && d_cnt_len && d_span && d_reg && d_bits_off && d_bits, "null pointer");
    AMPIS_REQUIRE(((uintptr_t)d_bits & 15u) == 0, "bits arena must be 16-byte aligned");
    const int grid = n < (1 << 20) ? n : (1 << 20);
    rle_paint_kernel<256><<<grid, 256, 0, as_stream(stream)>>>(
        d_cum, d_cnt_off, d_cnt_len, (const uint2 *)d_span, (const uint2 *)d_reg, d_bits_off, n, (uint4 *)d_bits,
        bits_capacity);
    AMPIS_CHECK_LAUNCH("rle_paint_kernel");
    return AMPIS_OK;
}

// ---- fused measure + paint ---------------------------------------------------------------------
// One launch instead of five (measure, 3 x scan, paint): a team of TEAM_WARPS warps owns one
// mask.  The team's first warp measures the mask (area, box, span) and leaves the run end
// positions in shared memory; the CTA then reserves arena space for all its masks with ONE
// atomicAdd on a global cursor (arena order is therefore arbitrary, which nothing depends on) and
// every team paints its region.  SPAN layout: TEAM_WARPS = 1 (8 masks per CTA, regions are a few
// KB); FULL layout: TEAM_WARPS = 8 (one mask per CTA, regions are 100s of KB).
#define MP_WARPS 8
#define MP_CUM_WORDS 2048      // shared run-end words per CTA, split evenly between its masks
#define MP_TILE 128            // chunks (2 KB) of packed mask a warp assembles in shared memory at a time

// first index r in [0,m) with C[r] > b (C ascending); m if none
__device__ __forceinline__ int upper_bound_u32(const u32 *C, int m, u64 b)
{
    int lo = 0, hi = m;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((u64)C[mid] > b) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// One warp assembles chunks [t0,t1) (at most MP_TILE) of a mask in its shared tile and streams
// them out.  Run driven: the tile is zeroed, then every lane takes 1-runs that intersect the
// tile and sets their bits (whole words by plain stores, the two boundary words by shared
// atomicOr since neighbouring runs may share a word).  A typical particle has ~40 1-runs, i.e.
// about one per lane -- far less work than locating each chunk by binary search.
__device__ __forceinline__ void warp_paint_tile(const u32 *C, int m, u32 t0, u32 t1, u32 *tile, uint4 *out,
                                                u32 lane)
{
    const u32 nch = t1 - t0;
    uint4 *tile4 = reinterpret_cast<uint4 *>(tile);
    for (u32 k = lane; k < nch; k += 32) tile4[k] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    const u64 b0 = (u64)t0 * AMPIS_CHUNK_BITS, b1 = (u64)t1 * AMPIS_CHUNK_BITS;
    const int r0 = upper_bound_u32(C, m, b0);          // run that owns bit b0
    const int r1 = upper_bound_u32(C, m, b1 - 1);      // run that owns bit b1-1 (m if beyond the runs)
    for (int r = (r0 | 1) + 2 * (int)lane; r <= r1 && r < m; r += 64) {   // odd runs are the 1-runs
        const u64 rs = (u64)C[r - 1], re = (u64)C[r];
        const u64 s64 = rs > b0 ? rs : b0, e64 = re < b1 ? re : b1;
        if (e64 <= s64) continue;
        const u32 s = (u32)(s64 - b0), e = (u32)(e64 - b0);      // bit range inside the tile
        const u32 w0 = s >> 5, w1 = (e - 1) >> 5;
        if (w0 == w1) {
            atomicOr(&tile[w0], bit_range(s & 31u, ((e - 1) & 31u) + 1u));
        } else {
            atomicOr(&tile[w0], bit_range(s & 31u, 32u));
            for (u32 w = w0 + 1; w < w1; w++) tile[w] = 0xffffffffu;
            atomicOr(&tile[w1], bit_range(0u, ((e - 1) & 31u) + 1u));
        }
    }
    __syncwarp();
    for (u32 k = lane; k < nch; k += 32) st_v4_stream(out + t0 + k, tile4[k]);
    __syncwarp();
}

// Same, for a mask whose run ends all sit in shared memory (m <= cap): no search at all -- every
// lane walks the 1-runs it owns (r = 1 + 2*lane, +64, ...) and clips them to the tile.
__device__ __forceinline__ void warp_paint_tile_smem(const u32 *sC, int m, u32 t0, u32 t1, u32 *tile, uint4 *out,
                                                     u32 lane)
{
    const u32 nch = t1 - t0;
    uint4 *tile4 = reinterpret_cast<uint4 *>(tile);
    for (u32 k = lane; k < nch; k += 32) tile4[k] = make_uint4(0u, 0u, 0u, 0u);
    __syncwarp();
    const u64 b0 = (u64)t0 * AMPIS_CHUNK_BITS, b1 = (u64)t1 * AMPIS_CHUNK_BITS;
    for (int r = 1 + 2 * (int)lane; r < m; r += 64) {
        const u64 rs = (u64)sC[r - 1], re = (u64)sC[r];
        if (rs >= b1) break;
        const u64 s64 = rs > b0 ? rs : b0, e64 = re < b1 ? re : b1;
        if (e64 <= s64) continue;
        const u32 s = (u32)(s64 - b0), e = (u32)(e64 - b0);
        const u32 w0 = s >> 5, w1 = (e - 1) >> 5;
        if (w0 == w1) {
            atomicOr(&tile[w0], bit_range(s & 31u, ((e - 1) & 31u) + 1u));
        } else {
            atomicOr(&tile[w0], bit_range(s & 31u, 32u));
            for (u32 w = w0 + 1; w < w1; w++) tile[w] = 0xffffffffu;
            atomicOr(&tile[w1], bit_range(0u, ((e - 1) & 31u) + 1u));
        }
    }
    __syncwarp();
    for (u32 k = lane; k < nch; k += 32) st_v4_stream(out + t0 + k, tile4[k]);
    __syncwarp();
}

// AMPIS_LAYOUT_CROP: one warp assembles the bounding-box window of a mask (columns bb.x..bb.z, for
// each column the absolute 32-row bands (bb.y>>5)..(bb.w>>5)) in shared-memory tiles of whole
// columns and writes the words out, coalesced.  Run driven like warp_paint_tile: each lane takes
// 1-runs, cuts them at column boundaries (a run that reaches the bottom of a column continues at
// the top of the next) and sets the bits of every piece.  `searched` = false when all runs are
// known to fall in one tile (small masks: no binary search at all).
template <int L = 32>
__device__ __forceinline__ void warp_paint_crop(const u32 *C, int m, u32 H, const int4 bb, u32 *tile,
                                                u32 tile_words, u32 *out, u32 lane)
{
    // `lane` = lane inside the group of L lanes that owns the mask (see warp_measure)
    const u32 gm = group_mask<L>();
    if (bb.z < bb.x) return;
    const u32 wy0 = (u32)bb.y >> 5, nwy = ((u32)bb.w >> 5) - wy0 + 1u;
    const u32 x_end = (u32)bb.z + 1u;
    const FastDiv byH = fastdiv_make(H);
    // band slices: one slice holds all bands of the box unless the box is taller than 32 * tile_words rows (then a
    // tile is a piece of ONE column and the window is assembled slice by slice)
    for (u32 wa = 0; wa < nwy; wa += tile_words) {
        const u32 nwb = min(tile_words, nwy - wa);
        const bool whole = nwb == nwy;
        const u32 y_lo = max((u32)bb.y, (wy0 + wa) << 5), y_hi = min((u32)bb.w + 1u, (wy0 + wa + nwb) << 5);
        const u32 cols_per_tile = whole ? max(1u, tile_words / nwb) : 1u;
        const bool one_tile = whole && (x_end - (u32)bb.x) <= cols_per_tile;
        for (u32 xa = (u32)bb.x; xa < x_end; xa += cols_per_tile) {
            const u32 xb = min(xa + cols_per_tile, x_end);
            const u32 tw = (xb - xa) * nwb;
            for (u32 k = lane; k < tw; k += L) tile[k] = 0u;
            __syncwarp(gm);
            const u64 b0 = (u64)xa * H, b1 = (u64)xb * H;
            int r0 = 0, r1 = m - 1;
            if (!one_tile) {
                r0 = upper_bound_u32(C, m, b0);          // run that owns bit b0
                r1 = upper_bound_u32(C, m, b1 - 1);      // run that owns bit b1-1 (m if beyond the runs)
            }
            for (int r = (r0 | 1) + 2 * (int)lane; r <= r1 && r < m; r += 2 * L) {   // odd runs are the 1-runs
                const u64 rs = (u64)C[r - 1], re = (u64)C[r];
                u64 s = rs > b0 ? rs : b0;
                const u64 e = re < b1 ? re : b1;
                while (s < e) {
                    const u32 x = fastdiv((u32)s, byH);               // s < e <= a run end, which is a u32
                    const u64 cs = (u64)x * H;
                    // rows [ys,ye) of column x, clipped to the slice's rows of the box (a well-formed mask never
                    // needs the box clip; a malformed one -- flagged in status -- must not write outside its window)
                    const u32 ys = max((u32)(s - cs), y_lo), ye = min((u32)(min(e, cs + H) - cs), y_hi);
                    s = cs + H;
                    if (ye <= ys) continue;
                    u32 *col = tile + (x - xa) * nwb - (wy0 + wa);
                    const u32 w0 = ys >> 5, w1 = (ye - 1) >> 5;
                    if (w0 == w1) {
                        atomicOr(&col[w0], bit_range(ys & 31u, ((ye - 1) & 31u) + 1u));
                    } else {
                        atomicOr(&col[w0], bit_range(ys & 31u, 32u));
                        for (u32 w = w0 + 1; w < w1; w++) col[w] = 0xffffffffu;
                        atomicOr(&col[w1], bit_range(0u, ((ye - 1) & 31u) + 1u));
                    }
                }
            }
            __syncwarp(gm);
            if (whole) {
                u32 *o = out + (xa - (u32)bb.x) * nwy;
                for (u32 k = lane; k < tw; k += L) o[k] = tile[k];
            } else {                                          // a piece of one column
                u32 *o = out + (xa - (u32)bb.x) * nwy + wa;
                for (u32 k = lane; k < tw; k += L) o[k] = tile[k];
            }
            __syncwarp(gm);
        }
    }
}

// KIND 0: in-span chunks located by binary search (FULL layout, where 97 % of the stores are zeros
// outside the span and occupancy of the store stream matters most); KIND 1: the span is assembled
// run-driven in shared-memory tiles (SPAN layout, where the span is the whole job); KIND 2: the
// bounding-box window is assembled the same way (CROP layout).
template <int TEAM_WARPS, int KIND>
__global__ void __launch_bounds__(MP_WARPS * 32, KIND ? 5 : 8)
rle_measure_paint_kernel(const u32 *__restrict__ cnt, const i64 *__restrict__ cnt_off,
                         const int *__restrict__ cnt_len, const u32 *__restrict__ hh,
                         const u32 *__restrict__ ww, int n, int layout, u32 *cum_g, u32 *__restrict__ area,
                         int *__restrict__ bbox, u32 *__restrict__ span, u32 *__restrict__ reg,
                         i64 *__restrict__ bits_off, int *__restrict__ status, uint4 *__restrict__ bits,
                         i64 capacity, unsigned long long *__restrict__ cursor)
{
    constexpr int MASKS = MP_WARPS / TEAM_WARPS;
    constexpr int CUM_CAP = MP_CUM_WORDS / MASKS;
    __shared__ __align__(16) u32 s_cum[MP_CUM_WORDS];
    constexpr bool TILED = KIND == 1;
    __shared__ __align__(16) u32 s_tile[KIND ? MP_WARPS : 1][KIND ? MP_TILE * 4 : 4];
    __shared__ uint2 s_span[MASKS], s_reg[MASKS];
    __shared__ int4 s_bbox[MASKS];
    __shared__ i64 s_off[MASKS];
    __shared__ i64 s_base;
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const int team = (int)wid / TEAM_WARPS, tw = (int)wid % TEAM_WARPS;
    const int i = blockIdx.x * MASKS + team;
    const bool valid = i < n;
    int m = 0;
    i64 base = 0;
    if (valid) { m = cnt_len[i]; base = cnt_off[i]; }
    if (tw == 0) {
        uint2 sp = make_uint2(0u, 0u), rg = sp;
        if (valid) {
            const u32 H = hh[i];
            const u64 HW = (u64)H * ww[i];
            const MaskMeasure ms = warp_measure(cnt + base, m, H, HW, s_cum + team * CUM_CAP, CUM_CAP,
                                                m > CUM_CAP ? cum_g + base : nullptr);
            if (lane == 0) store_measure(ms, H, HW, layout, i, area, bbox, span, reg, status, &sp, &rg, &s_bbox[team]);
        }
        if (lane == 0) { s_span[team] = sp; s_reg[team] = rg; }
    }
    i64 off;
    if (TEAM_WARPS == 1) {
        // a warp is its own team: it reserves its region with its own atomicAdd, no CTA-wide barrier
        // (the two __syncthreads of the shared reservation were 21 % of this kernel's stall samples)
        __syncwarp();
        if (!valid) return;
        unsigned long long o = 0;
        if (lane == 0) {
            const u32 sz = s_reg[team].y - s_reg[team].x;
            o = sz ? atomicAdd(cursor, (unsigned long long)sz) : 0ull;
        }
        off = (i64)__shfl_sync(0xffffffffu, o, 0);
    } else {
        __syncthreads();
        if (threadIdx.x == 0) {
            i64 tot = 0;
            for (int k = 0; k < MASKS; k++) { s_off[k] = tot; tot += (i64)(s_reg[k].y - s_reg[k].x); }
            s_base = tot ? (i64)atomicAdd(cursor, (unsigned long long)tot) : 0;
        }
        __syncthreads();
        if (!valid) return;
        off = s_base + s_off[team];
    }
    const uint2 rg = s_reg[team], sp = s_span[team];
    if (off + (i64)(rg.y - rg.x) > capacity) {
        // arena exhausted: leave the mask empty so later kernels stay inside the arena; the
        // caller sees *cursor > capacity and retries (ampis_rle_measure_paint contract)
        if (tw == 0 && lane == 0) {
            bits_off[i] = 0;
            reinterpret_cast<uint2 *>(span)[i] = make_uint2(0u, 0u);
            reinterpret_cast<uint2 *>(reg)[i] = make_uint2(0u, 0u);
            // the crop rows kernels size their reads from the box, not from reg
            if (KIND == 2) reinterpret_cast<int4 *>(bbox)[i] = make_int4(0, 0, -1, -1);
        }
        return;
    }
    if (tw == 0 && lane == 0) bits_off[i] = off;
    const u32 *C = m > CUM_CAP ? cum_g + base : s_cum + team * CUM_CAP;
    if (KIND == 2) {
        // bounding-box window, one warp per mask
        warp_paint_crop(C, m, hh[i], s_bbox[team], s_tile[KIND ? wid : 0], MP_TILE * 4,
                        reinterpret_cast<u32 *>(bits + off), lane);
        return;
    }
    uint4 *out = bits + off - rg.x;
    // zeros outside the span (FULL layout only; in SPAN layout region == span)
    for (u32 c = rg.x + tw * 32 + lane; c < sp.x; c += TEAM_WARPS * 32) st_v4_stream(out + c, make_uint4(0u, 0u, 0u, 0u));
    for (u32 c = max(sp.y, rg.x) + tw * 32 + lane; c < rg.y; c += TEAM_WARPS * 32)
        st_v4_stream(out + c, make_uint4(0u, 0u, 0u, 0u));
    if (TILED) {
        // the span itself, one tile per warp at a time
        if (m <= CUM_CAP) {
            for (u32 t0 = sp.x + (u32)tw * MP_TILE; t0 < sp.y; t0 += TEAM_WARPS * MP_TILE)
                warp_paint_tile_smem(s_cum + team * CUM_CAP, m, t0, min(t0 + MP_TILE, sp.y),
                                     s_tile[TILED ? wid : 0], out, lane);
        } else {
            for (u32 t0 = sp.x + (u32)tw * MP_TILE; t0 < sp.y; t0 += TEAM_WARPS * MP_TILE)
                warp_paint_tile(C, m, t0, min(t0 + MP_TILE, sp.y), s_tile[TILED ? wid : 0], out, lane);
        }
    } else {
        for (u32 c = sp.x + tw * 32 + lane; c < sp.y; c += TEAM_WARPS * 32)
            st_v4_stream(out + c, paint_chunk<false>(C, m, c));
    }
}

// Fused measure + paint for AMPIS_LAYOUT_CROP with a GROUP of L = 8 or 16 lanes per mask (four or two masks per
// warp): small instances (spheroidite carbides ~30 runs, powder particles ~75) leave most of a warp idle in the
// warp-per-mask kernel above, whose ~850 instructions per mask are mostly straight-line code paid per warp.
// Same steps: the group measures its mask (4 runs per lane and pass), lane 0 reserves arena space with its own
// atomicAdd, the group paints the bounding-box window through its slice of the shared tile.
template <int L>
__global__ void __launch_bounds__(MP_WARPS * 32, 6)
rle_measure_paint_crop_kernel(const u32 *__restrict__ cnt, const i64 *__restrict__ cnt_off,
                              const int *__restrict__ cnt_len, const u32 *__restrict__ hh,
                              const u32 *__restrict__ ww, int n, u32 *cum_g, u32 *__restrict__ area,
                              int *__restrict__ bbox, u32 *__restrict__ span, u32 *__restrict__ reg,
                              i64 *__restrict__ bits_off, int *__restrict__ status, uint4 *__restrict__ bits,
                              i64 capacity, unsigned long long *__restrict__ cursor)
{
    constexpr int GROUPS = MP_WARPS * 32 / L;
    constexpr int CUM_CAP = MP_CUM_WORDS / GROUPS;
    constexpr int TILE_WORDS = MP_WARPS * MP_TILE * 4 / GROUPS;
    __shared__ __align__(16) u32 s_cum[MP_CUM_WORDS];
    __shared__ __align__(16) u32 s_tile[MP_WARPS * MP_TILE * 4];
    __shared__ uint2 s_span[GROUPS], s_reg[GROUPS];
    __shared__ int4 s_bbox[GROUPS];
    const u32 gl = threadIdx.x & (u32)(L - 1);
    const int grp = (int)threadIdx.x / L;
    const u32 gm = group_mask<L>();
    const int i = blockIdx.x * GROUPS + grp;
    const bool valid = i < n;
    int m = 0;
    i64 base = 0;
    u32 H = 1;
    uint2 rg = make_uint2(0u, 0u);
    if (valid) {
        m = cnt_len[i];
        base = cnt_off[i];
        H = hh[i];
        const u64 HW = (u64)H * ww[i];
        const MaskMeasure ms = warp_measure<L>(cnt + base, m, H, HW, s_cum + grp * CUM_CAP, CUM_CAP,
                                               m > CUM_CAP ? cum_g + base : nullptr);
        if (gl == 0)
            store_measure(ms, H, HW, AMPIS_LAYOUT_CROP, i, area, bbox, span, reg, status, &s_span[grp], &s_reg[grp],
                          &s_bbox[grp]);
        __syncwarp(gm);
        rg = s_reg[grp];
    }
    // arena space: ONE atomicAdd per warp for its 32 / L masks (a single cursor takes ~1 G atomics/s, which is what
    // bounded this kernel with an atomic per mask)
    __syncwarp();
    const u32 sz = (valid && gl == 0) ? rg.y - rg.x : 0u;
    u32 before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 32 / L; k++) {
        const u32 v = __shfl_sync(0xffffffffu, sz, k * L);
        if (k < (int)(lane_id() / L)) before += v;
        total += v;
    }
    unsigned long long o = 0;
    if (lane_id() == 0 && total) o = atomicAdd(cursor, (unsigned long long)total);
    const i64 off = (i64)__shfl_sync(0xffffffffu, o, 0) + (i64)before;
    if (!valid) return;
    if (off + (i64)(rg.y - rg.x) > capacity) {           // arena exhausted: see rle_measure_paint_kernel
        if (gl == 0) {
            bits_off[i] = 0;
            reinterpret_cast<uint2 *>(span)[i] = make_uint2(0u, 0u);
            reinterpret_cast<uint2 *>(reg)[i] = make_uint2(0u, 0u);
            reinterpret_cast<int4 *>(bbox)[i] = make_int4(0, 0, -1, -1);    // rows kernels size their reads from it
        }
        return;
    }
    if (gl == 0) bits_off[i] = off;
    const u32 *C = m > CUM_CAP ? cum_g + base : s_cum + grp * CUM_CAP;
    warp_paint_crop<L>(C, m, H, s_bbox[grp], s_tile + grp * TILE_WORDS, TILE_WORDS,
                       reinterpret_cast<u32 *>(bits + off), gl);
}

// CROP layout painter for tables measured by ampis_rle_measure (offsets from the scan): one warp
// per mask, run ends read from global memory.
__global__ void __launch_bounds__(256)
rle_paint_crop_kernel(const u32 *__restrict__ cum, const i64 *__restrict__ cnt_off,
                      const int *__restrict__ cnt_len, const int4 *__restrict__ bbox,
                      const u32 *__restrict__ hh, const i64 *__restrict__ bits_off, int n,
                      uint4 *__restrict__ bits, i64 capacity)
{
    __shared__ u32 s_tile[8][MP_TILE * 4];
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const int i = blockIdx.x * 8 + (int)wid;
    if (i >= n) return;
    const int4 bb = bbox[i];
    const i64 off = bits_off[i];
    if (off + (i64)((crop_words(bb) + 3u) / 4u) > capacity) return;
    warp_paint_crop(cum + cnt_off[i], cnt_len[i], hh[i], bb, s_tile[wid], MP_TILE * 4,
                    reinterpret_cast<u32 *>(bits + off), lane);
}

extern "C" int ampis_rle_decode_crop(const uint32_t *d_cum, const int64_t *d_cnt_off, const int32_t *d_cnt_len,
                                     const int32_t *d_bbox, const uint32_t *d_h, const int64_t *d_bits_off,
                                     int32_t n, void *d_bits, int64_t bits_capacity, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_cum && d_cnt_off && d_cnt_len && d_bbox && d_h && d_bits_off && d_bits, "null pointer");
    AMPIS_REQUIRE(((uintptr_t)d_bits & 15u) == 0, "bits arena must be 16-byte aligned");
    rle_paint_crop_kernel<<<(n + 7) / 8, 256, 0, as_stream(stream)>>>(
        d_cum, d_cnt_off, d_cnt_len, (const int4 *)d_bbox, d_h, d_bits_off, n, (uint4 *)d_bits, bits_capacity);
    AMPIS_CHECK_LAUNCH("rle_paint_crop_kernel");
    return AMPIS_OK;
}

// typical runs per mask (caller's hint) up to which 8 / 16 lanes work on one mask in the crop kernel
#define PAINT_L8_MAX_RUNS 40
#define PAINT_L16_MAX_RUNS 112

extern "C" int ampis_rle_measure_paint(const uint32_t *d_cnt, const int64_t *d_cnt_off,
                                       const int32_t *d_cnt_len, const uint32_t *d_h, const uint32_t *d_w,
                                       int32_t n, int32_t layout, uint32_t *d_cum, uint32_t *d_area,
                                       int32_t *d_bbox, uint32_t *d_span, uint32_t *d_reg, int64_t *d_bits_off,
                                       int32_t *d_status, void *d_bits, int64_t bits_capacity,
                                       uint64_t *d_cursor, int32_t runs_hint, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    AMPIS_REQUIRE(layout == AMPIS_LAYOUT_SPAN || layout == AMPIS_LAYOUT_FULL || layout == AMPIS_LAYOUT_CROP,
                  "bad layout");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_cnt && d_cnt_off && d_cnt_len && d_h && d_w && d_cum && d_area && d_bbox && d_span &&
                      d_reg && d_bits_off && d_status && d_bits && d_cursor, "null pointer");
    AMPIS_REQUIRE(((uintptr_t)d_bits & 15u) == 0, "bits arena must be 16-byte aligned");
    cudaError_t e = cudaMemsetAsync(d_cursor, 0, sizeof(uint64_t), as_stream(stream));
    if (e != cudaSuccess) { ampis_set_error("cursor memset: %s", cudaGetErrorString(e)); return AMPIS_ECUDA; }
    if (layout == AMPIS_LAYOUT_FULL)
        rle_measure_paint_kernel<MP_WARPS, 0><<<n, MP_WARPS * 32, 0, as_stream(stream)>>>(
            d_cnt, d_cnt_off, d_cnt_len, d_h, d_w, n, layout, d_cum, d_area, d_bbox, d_span, d_reg, d_bits_off,
            d_status, (uint4 *)d_bits, bits_capacity, (unsigned long long *)d_cursor);
    else if (layout == AMPIS_LAYOUT_CROP && runs_hint > 0 && runs_hint <= PAINT_L8_MAX_RUNS)
        rle_measure_paint_crop_kernel<8><<<(n + 31) / 32, MP_WARPS * 32, 0, as_stream(stream)>>>(
            d_cnt, d_cnt_off, d_cnt_len, d_h, d_w, n, d_cum, d_area, d_bbox, d_span, d_reg, d_bits_off, d_status,
            (uint4 *)d_bits, bits_capacity, (unsigned long long *)d_cursor);
    else if (layout == AMPIS_LAYOUT_CROP && runs_hint > 0 && runs_hint <= PAINT_L16_MAX_RUNS)
        rle_measure_paint_crop_kernel<16><<<(n + 15) / 16, MP_WARPS * 32, 0, as_stream(stream)>>>(
            d_cnt, d_cnt_off, d_cnt_len, d_h, d_w, n, d_cum, d_area, d_bbox, d_span, d_reg, d_bits_off, d_status,
            (uint4 *)d_bits, bits_capacity, (unsigned long long *)d_cursor);
    else if (layout == AMPIS_LAYOUT_CROP)
        rle_measure_paint_kernel<1, 2><<<(n + MP_WARPS - 1) / MP_WARPS, MP_WARPS * 32, 0, as_stream(stream)>>>(
            d_cnt, d_cnt_off, d_cnt_len, d_h, d_w, n, layout, d_cum, d_area, d_bbox, d_span, d_reg, d_bits_off,
            d_status, (uint4 *)d_bits, bits_capacity, (unsigned long long *)d_cursor);
    else
        rle_measure_paint_kernel<1, 1><<<(n + MP_WARPS - 1) / MP_WARPS, MP_WARPS * 32, 0, as_stream(stream)>>>(
            d_cnt, d_cnt_off, d_cnt_len, d_h, d_w, n, layout, d_cum, d_area, d_bbox, d_span, d_reg, d_bits_off,
            d_status, (uint4 *)d_bits, bits_capacity, (unsigned long long *)d_cursor);
    AMPIS_CHECK_LAUNCH("rle_measure_paint_kernel");
    return AMPIS_OK;
}

// Fallback of the flat decode kernel (rle_flat.cu): masks it left on a list -- too many runs or too large a window
// for its per-warp budgets, or a frame of 2^31 pixels or more -- are measured and painted here, one warp per mask,
// any size (run ends beyond the shared budget go through cum_g, the window is assembled tile by tile).  The list
// length is only known on the device: list[0] = number of masks, list[1..] = mask ids; warps stride over it.
__global__ void __launch_bounds__(MP_WARPS * 32, 5)
rle_measure_paint_list_kernel(const u32 *__restrict__ cnt, const i64 *__restrict__ cnt_off,
                              const int *__restrict__ cnt_len, const u32 *__restrict__ hh,
                              const u32 *__restrict__ ww, const int *__restrict__ list, u32 *cum_g,
                              u32 *__restrict__ area, int *__restrict__ bbox, u32 *__restrict__ span,
                              u32 *__restrict__ reg, i64 *__restrict__ bits_off, int *__restrict__ status,
                              uint4 *__restrict__ bits, i64 capacity, unsigned long long *__restrict__ cursor)
{
    constexpr int CUM_CAP = MP_CUM_WORDS / MP_WARPS;
    __shared__ __align__(16) u32 s_cum[MP_CUM_WORDS];
    __shared__ __align__(16) u32 s_tile[MP_WARPS][MP_TILE * 4];
    __shared__ uint2 s_span[MP_WARPS], s_reg[MP_WARPS];
    __shared__ int4 s_bbox[MP_WARPS];
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const int count = list[0];
    for (int q = blockIdx.x * MP_WARPS + (int)wid; q < count; q += gridDim.x * MP_WARPS) {
        const int i = list[1 + q];
        const int m = cnt_len[i];
        const i64 base = cnt_off[i];
        const u32 H = hh[i];
        const u64 HW = (u64)H * ww[i];
        __syncwarp();
        const MaskMeasure ms = warp_measure(cnt + base, m, H, HW, s_cum + wid * CUM_CAP, CUM_CAP,
                                            m > CUM_CAP ? cum_g + base : nullptr);
        if (lane == 0)
            store_measure(ms, H, HW, AMPIS_LAYOUT_CROP, i, area, bbox, span, reg, status, &s_span[wid], &s_reg[wid],
                          &s_bbox[wid]);
        __syncwarp();
        const u32 sz = s_reg[wid].y - s_reg[wid].x;
        unsigned long long o = 0;
        if (lane == 0 && sz) o = atomicAdd(cursor, (unsigned long long)sz);
        const i64 off = (i64)__shfl_sync(0xffffffffu, o, 0);
        if (off + (i64)sz > capacity) {                      // arena exhausted: see rle_flat_crop_kernel
            if (lane == 0) {
                bits_off[i] = 0;
                reinterpret_cast<int4 *>(bbox)[i] = make_int4(0, 0, -1, -1);
                reinterpret_cast<uint2 *>(span)[i] = make_uint2(0u, 0u);
                reinterpret_cast<uint2 *>(reg)[i] = make_uint2(0u, 0u);
            }
            continue;
        }
        if (lane == 0) bits_off[i] = off;
        __threadfence_block();                               // run ends written to cum_g are read back below
        const u32 *C = m > CUM_CAP ? cum_g + base : s_cum + wid * CUM_CAP;
        warp_paint_crop(C, m, H, s_bbox[wid], s_tile[wid], MP_TILE * 4, reinterpret_cast<u32 *>(bits + off), lane);
    }
}

int ampis_launch_measure_paint_list(const uint32_t *d_cnt, const int64_t *d_cnt_off, const int32_t *d_cnt_len,
                                    const uint32_t *d_h, const uint32_t *d_w, const int32_t *d_list, int32_t max_n,
                                    uint32_t *d_cum, uint32_t *d_area, int32_t *d_bbox, uint32_t *d_span,
                                    uint32_t *d_reg, int64_t *d_bits_off, int32_t *d_status, void *d_bits,
                                    int64_t bits_capacity, uint64_t *d_cursor, cudaStream_t st)
{
    // the list is usually empty or short: a few waves of warps that stride over it
    const int want = (max_n + MP_WARPS - 1) / MP_WARPS;
    const int grid = want < 148 * 4 ? want : 148 * 4;
    rle_measure_paint_list_kernel<<<grid, MP_WARPS * 32, 0, st>>>(
        d_cnt, d_cnt_off, d_cnt_len, d_h, d_w, d_list, d_cum, d_area, d_bbox, d_span, d_reg, d_bits_off, d_status,
        (uint4 *)d_bits, bits_capacity, (unsigned long long *)d_cursor);
    AMPIS_CHECK_LAUNCH("rle_measure_paint_list_kernel");
    return AMPIS_OK;
}

// ---- packed -> bool[n][h][w] ------------------------------------------------------------
// Output tile: 32 columns x 32 rows per warp step. Lane l loads the 32 pixels (x0+l, y0..y0+31)
// as one funnel-shifted word; the warp transposes through shuffles-free ballot: for each row
// yy the byte of column x0+l is bit yy of lane l's word, written as a 32-byte row segment.
__device__ __forceinline__ u32 load_bits32(const u32 *__restrict__ words, const uint2 rg, u64 bit, u64 nbits)
{
    // 32 bits starting at linear pixel `bit` (zero outside the stored region / beyond nbits)
    if (bit >= nbits) return 0u;
    const u64 wlo = (u64)rg.x * 4, whi = (u64)rg.y * 4;   // region in 32-bit words
    const u64 wi = bit >> 5;
    const u32 sh = (u32)(bit & 31);
    const u32 a = (wi >= wlo && wi < whi) ? words[wi - wlo] : 0u;
    const u32 b = (wi + 1 >= wlo && wi + 1 < whi) ? words[wi + 1 - wlo] : 0u;
    u32 v = __funnelshift_r(a, b, sh);
    const u64 left = nbits - bit;
    if (left < 32) v &= (1u << left) - 1u;
    return v;
}

__global__ void __launch_bounds__(256)
unpack_bool_nrc_kernel(const uint4 *__restrict__ bits, const i64 *__restrict__ bits_off,
                       const uint2 *__restrict__ reg, const int *__restrict__ ids, u32 h, u32 w,
                       uint8_t *__restrict__ out)
{
    const int k = blockIdx.z;
    const int id = ids ? ids[k] : k;
    const uint2 rg = reg[id];
    const u32 *words = reinterpret_cast<const u32 *>(bits + bits_off[id]);
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const u32 x = blockIdx.x * 32 + lane;
    const u32 y0 = (blockIdx.y * 8 + wid) * 32;
    if (y0 >= h) return;
    u32 v = 0;
    if (x < w) v = load_bits32(words, rg, (u64)x * h + y0, (u64)(x + 1) * h);
    uint8_t *o = out + (u64)k * h * w;
    const u32 rows = min(32u, h - y0);
    for (u32 yy = 0; yy < rows; yy++)
        if (x < w) o[(u64)(y0 + yy) * w + x] = (uint8_t)((v >> yy) & 1u);
}

extern "C" int ampis_unpack_bool_nrc(const void *d_bits, const int64_t *d_bits_off, const uint32_t *d_reg,
                                     const int32_t *d_mask_ids, int32_t n, uint32_t h, uint32_t w,
                                     uint8_t *d_out, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0 || h == 0 || w == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits && d_bits_off && d_reg && d_out, "null pointer");
    AMPIS_REQUIRE(n <= 65535, "at most 65535 masks per call");
    dim3 grid((w + 31) / 32, (h + 255) / 256, n);
    AMPIS_REQUIRE(grid.y <= 65535, "image too tall");
    unpack_bool_nrc_kernel<<<grid, 256, 0, as_stream(stream)>>>((const uint4 *)d_bits, d_bits_off,
                                                               (const uint2 *)d_reg, d_mask_ids, h, w, d_out);
    AMPIS_CHECK_LAUNCH("unpack_bool_nrc_kernel");
    return AMPIS_OK;
}

// ---- bool[n][h][w] -> packed (FULL layout) -------------------------------------------------
// Thread per output 32-bit word: gathers 32 pixels of one or two columns.
template <bool Y_MAJOR>
__global__ void __launch_bounds__(256)
pack_bool_nrc_kernel(const uint8_t *__restrict__ masks, u32 h, u32 w, uint4 *__restrict__ bits,
                     const i64 *__restrict__ bits_off)
{
    const int k = blockIdx.y;
    const u64 hw = (u64)h * w;
    const u64 nwords = ((hw + 127) / 128) * 4;
    const u64 wi = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= nwords) return;
    const uint8_t *m = masks + (u64)k * hw;
    u32 v = 0;
    u64 p = wi * 32;
    u32 x = (u32)(p / h), y = (u32)(p - (u64)x * h);
#pragma unroll 4
    for (int b = 0; b < 32; b++) {
        // Y_MAJOR: the array is stored [n][w][h] (a transposed view of a Fortran-ordered stack, what
        // RLE.decode returns), i.e. already in COCO pixel order
        if (p + b < hw && (Y_MAJOR ? m[p + b] : m[(u64)y * w + x])) v |= 1u << b;
        if (++y == h) { y = 0; x++; }
    }
    reinterpret_cast<u32 *>(bits + bits_off[k])[wi] = v;
}

extern "C" int ampis_pack_bool_nrc(const uint8_t *d_masks, int32_t n, uint32_t h, uint32_t w, int32_t y_major,
                                   void *d_bits, const int64_t *d_bits_off, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0 || h == 0 || w == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_masks && d_bits && d_bits_off, "null pointer");
    AMPIS_REQUIRE(n <= 65535, "at most 65535 masks per call");
    const u64 nwords = (((u64)h * w + 127) / 128) * 4;
    dim3 grid((unsigned)((nwords + 255) / 256), n);
    if (y_major)
        pack_bool_nrc_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(d_masks, h, w, (uint4 *)d_bits, d_bits_off);
    else
        pack_bool_nrc_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(d_masks, h, w, (uint4 *)d_bits, d_bits_off);
    AMPIS_CHECK_LAUNCH("pack_bool_nrc_kernel");
    return AMPIS_OK;
}

// ---- bool[n][h][w] -> area + tight bbox ------------------------------------------------------
// The array is [n][h][w] (x fastest) or, with y_major, [n][w][h] (y fastest: a transposed view of a
// Fortran-ordered stack).  Threads read 16 bytes at a time along the fastest axis when rows allow it.
__global__ void __launch_bounds__(256)
bool_area_bbox_kernel(const uint8_t *__restrict__ masks, u32 h, u32 w, int y_major, u64 *__restrict__ area,
                      int *__restrict__ bbox)
{
    const int k = blockIdx.x;
    const u64 hw = (u64)h * w;
    const uint8_t *m = masks + (u64)k * hw;
    const u32 inner = y_major ? h : w;                  // length of the fastest axis
    u32 a = 0, i0 = 0xffffffffu, o0 = 0xffffffffu, i1 = 0, o1 = 0;      // inner / outer coordinate ranges
    if ((inner & 15u) == 0 && (((uintptr_t)m) & 15u) == 0) {
        const uint4 *m4 = reinterpret_cast<const uint4 *>(m);
        for (u64 q = threadIdx.x; q < hw / 16; q += blockDim.x) {
            const uint4 v = m4[q];
            if ((v.x | v.y | v.z | v.w) == 0u) continue;
            const u64 p = q * 16;
            const u32 o = (u32)(p / inner), ib = (u32)(p - (u64)o * inner);
            const u32 wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; j++) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    if ((wv[j] >> (8 * b)) & 0xffu) {
                        const u32 i = ib + 4u * j + b;
                        a++;
                        i0 = min(i0, i); i1 = max(i1, i + 1);
                    }
                }
            }
            o0 = min(o0, o); o1 = max(o1, o + 1);
        }
    } else {
        for (u64 p = threadIdx.x; p < hw; p += blockDim.x) {
            if (m[p]) {
                const u32 o = (u32)(p / inner), i = (u32)(p - (u64)o * inner);
                a++;
                i0 = min(i0, i); i1 = max(i1, i + 1);
                o0 = min(o0, o); o1 = max(o1, o + 1);
            }
        }
    }
    u32 x0 = y_major ? o0 : i0, x1 = y_major ? o1 : i1, y0 = y_major ? i0 : o0, y1 = y_major ? i1 : o1;
    __shared__ u32 sa[8], sx0[8], sy0[8], sx1[8], sy1[8];
    a = warp_sum(a); x0 = warp_min(x0); y0 = warp_min(y0); x1 = warp_max(x1); y1 = warp_max(y1);
    const u32 wid = threadIdx.x >> 5;
    if (lane_id() == 0) { sa[wid] = a; sx0[wid] = x0; sy0[wid] = y0; sx1[wid] = x1; sy1[wid] = y1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        u64 A = 0;
        for (int q = 0; q < 8; q++) {
            A += sa[q]; x0 = min(x0, sx0[q]); y0 = min(y0, sy0[q]); x1 = max(x1, sx1[q]); y1 = max(y1, sy1[q]);
        }
        area[k] = A;
        int4 bb = A ? make_int4((int)x0, (int)y0, (int)x1 - 1, (int)y1 - 1) : make_int4(0, 0, -1, -1);
        reinterpret_cast<int4 *>(bbox)[k] = bb;
    }
}

extern "C" int ampis_bool_area_bbox(const uint8_t *d_masks, int32_t n, uint32_t h, uint32_t w, int32_t y_major,
                                    uint64_t *d_area, int32_t *d_bbox, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_masks && d_area && d_bbox, "null pointer");
    bool_area_bbox_kernel<<<n, 256, 0, as_stream(stream)>>>(d_masks, h, w, y_major, (u64 *)d_area, d_bbox);
    AMPIS_CHECK_LAUNCH("bool_area_bbox_kernel");
    return AMPIS_OK;
}
