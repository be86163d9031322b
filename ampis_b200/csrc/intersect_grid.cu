// Intersection rows over AMPIS_LAYOUT_CROP tables with the bounding-box pre-pruning done through a
// uniform grid over the image instead of a scan of every column (same outputs as
// intersect_rows_crop_kernel: analyze.py:149-164 / powder.py:80-86 semantics, see intersect.cu).
//
// rleIou's bbIou pre-pass (what the reference relies on) tests all G x P boxes.  On images with
// thousands of small instances (spheroidite: 5,000 x 5,000 per 2048 x 2048 frame) that scan is the
// whole cost: 25 M box tests per image for ~5,000 overlapping pairs.  Here the column masks of an image
// are binned once into a 32 x 32 grid of square cells (cell side 2^shift pixels, at least the mean box
// side of the image's columns so that a mask lands in ~1-4 cells); a row then only looks at the
// columns registered in the cells its own box touches: ~15 box tests per row instead of 5,000.
//
//   grid_build_kernel   CTA per image: cell shift from the column boxes, per-cell counts and their scan in shared
//       memory, one atomicAdd on a global cursor for the image's entries, then every column mask writes its
//       index and its box into the cells it touches
//   intersect_rows_grid_kernel   eight lanes per row, four rows per warp: the cells of the row's box are read by
//       one lane each, their entry lists are flattened by a scan over the tile, lanes test one entry each (a pair
//       is taken only in the cell holding the top-left corner of the two boxes' overlap, so it is seen once),
//       candidates are intersected one after another by the row's eight lanes; overlaps of 256 words or more
//       are parked and intersected by the whole warp afterwards.
// Entries inside a cell are in no particular order: every arg-max tie is broken explicitly on the
// column index, so the result does not depend on it.
// Optional sparse output: (row, column, intersection) triplets of the non-zero intersections appended
// through an atomic cursor -- the "bbox-pruned sparse IoU" form for images whose dense G x P matrix
// (100 MB at 5,000 x 5,000) is not wanted.
#include "common.cuh"
#include "crop_common.cuh"

#define GR_N 32                 // cells per axis
#define GR_CELLS (GR_N * GR_N)
#define GR_ROWS AMPIS_ROWS_PER_CTA               // rows per CTA (= ampis_rows_per_block())
#define GR_TILE 8               // lanes per row
#define GR_THREADS (GR_ROWS * GR_TILE)
#define GR_LIST 16              // candidates a tile collects before it intersects them
#define GR_BIGLIST 16           // large overlaps a warp parks for its warp-wide phase
#define GR_BIG 256              // overlap words from which a candidate gets the whole warp

__device__ __forceinline__ int bits_of(u32 x) { return 32 - __clz(x); }
__device__ __forceinline__ int cell_of(int v, int shift) { return min(v >> shift, GR_N - 1); }

// One CTA builds the grid of one image: cell shift from the column boxes, per-cell counts in shared memory,
// exclusive scan, ONE atomicAdd on the global cursor for the image's entries, then the entries themselves
// (shared-memory cursors per cell).  cell_off has GR_CELLS + 1 values per image (absolute entry positions).
#define GB_THREADS 512

__global__ void __launch_bounds__(GB_THREADS)
grid_build_kernel(const int4 *__restrict__ bbox, const int *__restrict__ grp_col_begin,
                  const int *__restrict__ grp_col_count, int *__restrict__ grp_shift, i64 *__restrict__ cell_off,
                  int *__restrict__ entries, int4 *__restrict__ entry_bbox, i64 capacity,
                  unsigned long long *__restrict__ cursor)
{
    __shared__ u32 s_cnt[GR_CELLS], s_off[GR_CELLS];
    __shared__ u32 s_wsum[GB_THREADS / 32];
    __shared__ unsigned long long s_sum;
    __shared__ u32 s_n, s_ext, s_total;
    __shared__ int s_shift;
    __shared__ i64 s_base;
    const int g = blockIdx.x;
    const u32 tid = threadIdx.x, lane = lane_id(), wid = tid >> 5;
    const int cb = grp_col_begin[g], P = grp_col_count[g];
    if (tid == 0) { s_sum = 0; s_n = 0; s_ext = 0; }
    for (int c = tid; c < GR_CELLS; c += GB_THREADS) s_cnt[c] = 0;
    __syncthreads();
    // ---- cell size: 2^shift >= mean box side, and the image's extent fits 32 cells
    u32 sum = 0, n = 0, ext = 0;
    for (int k = tid; k < P; k += GB_THREADS) {
        const int4 b = bbox[cb + k];
        if (b.z < b.x) continue;
        sum += (u32)max(b.z - b.x, b.w - b.y) + 1u;
        ext = max(ext, (u32)max(b.z, b.w));
        n++;
    }
    sum = warp_sum(sum); n = warp_sum(n); ext = warp_max(ext);
    if (lane == 0) { atomicAdd(&s_sum, (unsigned long long)sum); atomicAdd(&s_n, n); atomicMax(&s_ext, ext); }
    __syncthreads();
    if (tid == 0) {
        const u32 mean = s_n ? (u32)((s_sum + s_n - 1) / s_n) : 1u;
        const int by_frame = max(0, bits_of(s_ext) - 5);               // extent >> shift < 32
        const int by_size = bits_of(max(mean, 1u) - 1u);                // 2^shift >= mean box side
        s_shift = min(max(by_frame, by_size), 30);
        grp_shift[g] = s_shift;
    }
    __syncthreads();
    const int s = s_shift;
    // ---- count
    for (int k = tid; k < P; k += GB_THREADS) {
        const int4 b = bbox[cb + k];
        if (b.z < b.x) continue;
        const int cx0 = cell_of(b.x, s), cx1 = cell_of(b.z, s), cy0 = cell_of(b.y, s), cy1 = cell_of(b.w, s);
        for (int cy = cy0; cy <= cy1; cy++)
            for (int cx = cx0; cx <= cx1; cx++) atomicAdd(&s_cnt[cy * GR_N + cx], 1u);
    }
    __syncthreads();
    // ---- exclusive scan of the 1024 counts: two cells per thread, warp scan, scan of the warp sums
    const u32 c0 = s_cnt[2 * tid], c1 = s_cnt[2 * tid + 1];
    u32 incl = c0 + c1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u32 v = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += v;
    }
    if (lane == 31) s_wsum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        u32 w = lane < GB_THREADS / 32 ? s_wsum[lane] : 0u, wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 v = __shfl_up_sync(0xffffffffu, wi, d);
            if ((int)lane >= d) wi += v;
        }
        if (lane < GB_THREADS / 32) s_wsum[lane] = wi - w;             // exclusive
        if (lane == 31) {
            s_total = wi;
            s_base = wi ? (i64)atomicAdd(cursor, (unsigned long long)wi) : 0;
        }
    }
    __syncthreads();
    const u32 excl = s_wsum[wid] + incl - (c0 + c1);
    s_off[2 * tid] = excl;
    s_off[2 * tid + 1] = excl + c0;
    s_cnt[2 * tid] = 0;                                                 // reused as the fill cursors
    s_cnt[2 * tid + 1] = 0;
    const i64 base = s_base;
    i64 *off = cell_off + (i64)g * (GR_CELLS + 1);
    off[2 * tid] = base + excl;
    off[2 * tid + 1] = base + excl + c0;
    if (tid == 0) off[GR_CELLS] = base + s_total;
    __syncthreads();
    // ---- fill
    for (int k = tid; k < P; k += GB_THREADS) {
        const int4 b = bbox[cb + k];
        if (b.z < b.x) continue;
        const int cx0 = cell_of(b.x, s), cx1 = cell_of(b.z, s), cy0 = cell_of(b.y, s), cy1 = cell_of(b.w, s);
        for (int cy = cy0; cy <= cy1; cy++)
            for (int cx = cx0; cx <= cx1; cx++) {
                const int cell = cy * GR_N + cx;
                const i64 pos = base + s_off[cell] + atomicAdd(&s_cnt[cell], 1u);
                if (pos < capacity) { entries[pos] = k; entry_bbox[pos] = b; }
            }
    }
}

struct GridRowArgs {
    const u32 *words;
    const i64 *bits_off;
    const int4 *bbox;
    const u32 *area;
    const int *row_mask;
    const int *blk_grp, *blk_row0;
    const int *grp_row_begin, *grp_row_count, *grp_col_begin, *grp_col_count;
    const int *grp_shift;
    const i64 *cell_off;
    const int *entries;
    const int4 *entry_bbox;
    i64 capacity;
    const i64 *grp_imat_off;
    int *imat;
    int *best_col;
    u32 *best_inter;
    double *best_score;
    int *coo_row, *coo_col;
    u32 *coo_inter;
    i64 coo_capacity;
    unsigned long long *coo_count;
};

// Eight lanes per row (a warp works on four rows at once): with a handful of cells, entries and candidates per
// row, a whole warp per row left most lanes idle and the kernel latency bound (ncu: 53 % of the stall samples on
// dependent loads at 39 % occupancy).  The four 8-lane tiles of a warp run independently (tile-masked shuffles,
// ballots and syncs); candidates with large overlaps are parked in a per-warp list and handled by all 32 lanes
// once every tile is through, so big masks still get the whole warp.
template <int MODE>
__global__ void __launch_bounds__(GR_THREADS, 20)
intersect_rows_grid_kernel(const GridRowArgs p)
{
    __shared__ int s_k[GR_ROWS][GR_LIST];
    __shared__ int4 s_bb[GR_ROWS][GR_LIST];
    __shared__ int s_pre[GR_ROWS][GR_TILE];         // exclusive prefix of the entry counts of the tile's 8 cells
    __shared__ int s_base[GR_ROWS][GR_TILE];        // first entry of the cell (relative to the image) - prefix
    __shared__ int s_nbig[GR_THREADS / 32];
    __shared__ int s_big_k[GR_THREADS / 32][GR_BIGLIST], s_big_slot[GR_THREADS / 32][GR_BIGLIST];
    __shared__ u32 s_big_area[GR_THREADS / 32][GR_BIGLIST];
    __shared__ i64 s_big_off[GR_THREADS / 32][GR_BIGLIST];
    __shared__ int4 s_big_bb[GR_THREADS / 32][GR_BIGLIST];
    __shared__ int4 s_row_bb[GR_ROWS];
    __shared__ const u32 *s_row_A[GR_ROWS];

    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const u32 t = lane & (GR_TILE - 1), tsh = lane & ~(GR_TILE - 1);
    const u32 tmask = ((1u << GR_TILE) - 1u) << tsh;
    const int slot = (int)threadIdx.x / GR_TILE;
    const int g = p.blk_grp[blockIdx.x];
    const int r = p.blk_row0[blockIdx.x] + slot;
    const bool valid = r < p.grp_row_begin[g] + p.grp_row_count[g];
    const int cb = p.grp_col_begin[g];
    const int P = p.grp_col_count[g];
    const i64 imat_off = (p.imat && p.grp_imat_off) ? p.grp_imat_off[g] : -1;
    int *irow = (valid && imat_off >= 0) ? p.imat + imat_off + (i64)(r - p.grp_row_begin[g]) * P : nullptr;

    int4 rb = make_int4(0, 0, -1, -1);
    u32 ra = 0;
    const u32 *A = nullptr;
    if (valid) {
        const int rm = p.row_mask[r];
        rb = p.bbox[rm];
        ra = p.area[rm];
        A = p.words + p.bits_off[rm] * 4;
    }
    if (t == 0) { s_row_bb[slot] = rb; s_row_A[slot] = A; }
    if (lane == 0) s_nbig[wid] = 0;
    __syncwarp();
    double best_s = 0.0;                             // identical in the 8 lanes of a tile
    u32 best_i = 0;
    int best_c = MODE == AMPIS_MODE_IOU ? -1 : (P > 0 ? 0 : -1);

    auto update = [&](int k, u32 inter, u32 ca) {
        if (!inter) return;
        if (t == 0) {
            if (irow) irow[k] = (int)inter;
            if (p.coo_count) {
                const unsigned long long pos = atomicAdd(p.coo_count, 1ull);
                if ((i64)pos < p.coo_capacity) { p.coo_row[pos] = r; p.coo_col[pos] = k; p.coo_inter[pos] = inter; }
            }
        }
        if (MODE == AMPIS_MODE_IOU) {
            const double s = (double)inter / (double)(ra + ca - inter);
            if (s > best_s || (s == best_s && (unsigned)k < (unsigned)best_c)) { best_s = s; best_i = inter; best_c = k; }
        } else {
            if (inter > best_i || (inter == best_i && (unsigned)k < (unsigned)best_c)) { best_i = inter; best_c = k; }
        }
    };

    // the tile's candidate list -> intersections; metadata of up to 8 candidates is fetched by one lane each
    auto flush = [&](int n) {
        for (int j0 = 0; j0 < n; j0 += GR_TILE) {
            const int mine = j0 + (int)t;
            int k_m = 0;
            i64 off_m = 0;
            u32 area_m = 0;
            if (mine < n) {
                k_m = s_k[slot][mine];
                off_m = p.bits_off[cb + k_m];
                area_m = p.area[cb + k_m];
            }
            const int nq = min(GR_TILE, n - j0);
            // two candidates at a time, four lanes each: their window words are in flight together (the kernel is
            // bound by the latency of these dependent loads, not by lanes)
            const u32 sub = t >> 2, t4 = t & 3u;
            for (int q0 = 0; q0 < nq; q0 += 2) {
                const int q = q0 + (int)sub;
                const bool have = q < nq;
                const int qs = have ? q : q0;
                const int k = __shfl_sync(tmask, k_m, qs, GR_TILE);
                const i64 off = __shfl_sync(tmask, off_m, qs, GR_TILE);
                const u32 ca = __shfl_sync(tmask, area_m, qs, GR_TILE);
                const int4 cbx = s_bb[slot][j0 + qs];
                const Overlap o = overlap_of(A, rb, p.words + off * 4, cbx);
                // large overlaps are parked for the whole warp (the four lanes of a candidate decide alike; the
                // shuffle is executed by all eight lanes whatever the two halves decide)
                const bool big = have && o.total >= GR_BIG;
                int pos = GR_BIGLIST;
                if (big && t4 == 0) pos = atomicAdd(&s_nbig[wid], 1);
                pos = __shfl_sync(tmask, pos, (int)(sub * 4u), GR_TILE);
                const bool parked = big && pos < GR_BIGLIST;
                if (parked && t4 == 0) {
                    s_big_k[wid][pos] = k; s_big_slot[wid][pos] = slot; s_big_area[wid][pos] = ca;
                    s_big_off[wid][pos] = off; s_big_bb[wid][pos] = cbx;
                }
                u32 v = (have && !parked) ? overlap_popc(o, t4, 4) : 0u;
                v += __shfl_xor_sync(tmask, v, 2);
                v += __shfl_xor_sync(tmask, v, 1);
                // both results to all eight lanes, applied in candidate order (every lane keeps the same best)
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const u32 vu = __shfl_sync(tmask, v, u * 4, GR_TILE);
                    const int ku = __shfl_sync(tmask, k, u * 4, GR_TILE);
                    const u32 cau = __shfl_sync(tmask, ca, u * 4, GR_TILE);
                    if (q0 + u < nq) update(ku, vu, cau);     // parked or empty: vu == 0, nothing happens
                }
            }
        }
    };

    if (valid && ra != 0 && P > 0) {
        const int s = p.grp_shift[g];
        const i64 *off = p.cell_off + (i64)g * (GR_CELLS + 1);
        const i64 gbase = off[0];
        const int rcx0 = cell_of(rb.x, s), rcy0 = cell_of(rb.y, s);
        const int ncx = cell_of(rb.z, s) - rcx0 + 1, ncell = ncx * (cell_of(rb.w, s) - rcy0 + 1);
        int n = 0;
        for (int c0 = 0; c0 < ncell; c0 += GR_TILE) {
            // one cell per lane: entry range, flattened by an exclusive scan of the lengths over the tile
            const int ci = c0 + (int)t;
            int st = 0, len = 0;
            if (ci < ncell) {
                const int cell = (rcy0 + ci / ncx) * GR_N + rcx0 + ci % ncx;
                st = (int)(off[cell] - gbase);
                len = (int)(off[cell + 1] - gbase) - st;
            }
            int incl = len;
#pragma unroll
            for (int d = 1; d < GR_TILE; d <<= 1) {
                const int v = __shfl_up_sync(tmask, incl, d, GR_TILE);
                if ((int)t >= d) incl += v;
            }
            const int T = __shfl_sync(tmask, incl, GR_TILE - 1, GR_TILE);
            __syncwarp(tmask);                                     // previous round's readers are done
            s_pre[slot][t] = incl - len;
            s_base[slot][t] = st - (incl - len);
            __syncwarp(tmask);
            for (int t0 = 0; t0 < T; t0 += GR_TILE) {
                const int idx = t0 + (int)t;
                bool cand = false;
                int k = 0;
                int4 b = make_int4(0, 0, -1, -1);
                if (idx < T) {
                    int j = 0;                                     // last cell with prefix <= idx (empty cells share a prefix)
#pragma unroll
                    for (int d = GR_TILE / 2; d; d >>= 1)
                        if (s_pre[slot][j + d] <= idx) j += d;
                    const i64 e = gbase + s_base[slot][j] + idx;
                    if (e < p.capacity) {
                        k = p.entries[e];
                        b = p.entry_bbox[e];
                        const int cj = c0 + j;
                        cand = b.x <= rb.z && b.z >= rb.x && b.y <= rb.w && b.w >= rb.y &&
                               max(rcx0, cell_of(b.x, s)) == rcx0 + cj % ncx &&
                               max(rcy0, cell_of(b.y, s)) == rcy0 + cj / ncx;
                    }
                }
                const u32 bal = (__ballot_sync(tmask, cand) >> tsh) & ((1u << GR_TILE) - 1u);
                if (cand) {
                    const int pos = n + __popc(bal & ((1u << t) - 1u));
                    s_k[slot][pos] = k;
                    s_bb[slot][pos] = b;
                }
                n += __popc(bal);
                if (n > GR_LIST - GR_TILE) {
                    __syncwarp(tmask);
                    flush(n);
                    n = 0;
                    __syncwarp(tmask);
                }
            }
        }
        __syncwarp(tmask);
        flush(n);
    }
    // ---- large overlaps parked by the warp's tiles: all 32 lanes on one candidate at a time
    __syncwarp();
    const int nbig = min(s_nbig[wid], GR_BIGLIST);
    for (int i = 0; i < nbig; i++) {
        const int os = s_big_slot[wid][i];
        const Overlap o = overlap_of(s_row_A[os], s_row_bb[os], p.words + s_big_off[wid][i] * 4, s_big_bb[wid][i]);
        const u32 v = warp_sum(overlap_popc(o, lane, 32));
        if (os == slot) update(s_big_k[wid][i], v, s_big_area[wid][i]);
    }
    if (valid && t == 0) {
        if (MODE == AMPIS_MODE_SAT) best_s = (double)best_i / (double)ra;   // 0/0 = NaN like numpy
        p.best_col[r] = best_c;
        p.best_inter[r] = best_i;
        p.best_score[r] = best_s;
    }
}

extern "C" int ampis_grid_cells(void) { return GR_CELLS; }

extern "C" int ampis_grid_build(const int32_t *d_bbox, const int32_t *d_grp_col_begin,
                                const int32_t *d_grp_col_count, int32_t n_groups, int32_t *d_grp_shift,
                                int64_t *d_cell_off, int32_t *d_entries, int32_t *d_entry_bbox, int64_t capacity,
                                uint64_t *d_cursor, void *stream)
{
    AMPIS_REQUIRE(n_groups >= 0 && capacity >= 0, "negative size");
    if (n_groups == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bbox && d_grp_col_begin && d_grp_col_count && d_grp_shift && d_cell_off && d_cursor,
                  "null pointer");
    AMPIS_REQUIRE((d_entries && d_entry_bbox) || capacity == 0, "entries missing");
    static_assert(GB_THREADS * 2 == GR_CELLS, "two cells per thread");
    cudaError_t e = cudaMemsetAsync(d_cursor, 0, sizeof(uint64_t), as_stream(stream));
    if (e != cudaSuccess) { ampis_set_error("cursor memset: %s", cudaGetErrorString(e)); return AMPIS_ECUDA; }
    grid_build_kernel<<<n_groups, GB_THREADS, 0, as_stream(stream)>>>(
        (const int4 *)d_bbox, d_grp_col_begin, d_grp_col_count, d_grp_shift, d_cell_off, d_entries,
        (int4 *)d_entry_bbox, capacity, (unsigned long long *)d_cursor);
    AMPIS_CHECK_LAUNCH("grid_build_kernel");
    return AMPIS_OK;
}

extern "C" int ampis_intersect_rows_grid(const void *d_bits, const int64_t *d_bits_off, const int32_t *d_bbox,
                                         const uint32_t *d_area, const int32_t *d_row_mask,
                                         const int32_t *d_blk_grp, const int32_t *d_blk_row0, int32_t n_blocks,
                                         const int32_t *d_grp_row_begin, const int32_t *d_grp_row_count,
                                         const int32_t *d_grp_col_begin, const int32_t *d_grp_col_count,
                                         const int32_t *d_grp_shift, const int64_t *d_cell_off,
                                         const int32_t *d_entries, const int32_t *d_entry_bbox, int64_t capacity,
                                         const int64_t *d_grp_imat_off, int32_t mode, int32_t *d_imat,
                                         int64_t imat_ints,
                                         int32_t *d_best_col, uint32_t *d_best_inter, double *d_best_score,
                                         int32_t *d_coo_row, int32_t *d_coo_col, uint32_t *d_coo_inter,
                                         int64_t coo_capacity, uint64_t *d_coo_count, void *stream)
{
    AMPIS_REQUIRE(n_blocks >= 0, "n_blocks < 0");
    AMPIS_REQUIRE(mode == AMPIS_MODE_IOU || mode == AMPIS_MODE_SAT, "bad mode");
    if (n_blocks == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits_off && d_bbox && d_area && d_row_mask && d_blk_grp && d_blk_row0 && d_grp_row_begin &&
                      d_grp_row_count && d_grp_col_begin && d_grp_col_count && d_grp_shift && d_cell_off &&
                      d_best_col && d_best_inter && d_best_score, "null pointer");
    AMPIS_REQUIRE((d_entries && d_entry_bbox) || capacity == 0, "entries missing");
    AMPIS_REQUIRE(imat_ints >= 0, "imat_ints < 0");
    AMPIS_REQUIRE(!d_coo_count || (d_coo_row && d_coo_col && d_coo_inter && coo_capacity >= 0) || coo_capacity == 0,
                  "sparse output arrays missing");
    GridRowArgs a;
    a.words = (const u32 *)d_bits; a.bits_off = d_bits_off; a.bbox = (const int4 *)d_bbox; a.area = d_area;
    a.row_mask = d_row_mask; a.blk_grp = d_blk_grp; a.blk_row0 = d_blk_row0;
    a.grp_row_begin = d_grp_row_begin; a.grp_row_count = d_grp_row_count;
    a.grp_col_begin = d_grp_col_begin; a.grp_col_count = d_grp_col_count;
    a.grp_shift = d_grp_shift; a.cell_off = d_cell_off; a.entries = d_entries; a.entry_bbox = (const int4 *)d_entry_bbox; a.capacity = capacity;
    a.grp_imat_off = d_grp_imat_off; a.imat = d_imat;
    a.best_col = d_best_col; a.best_inter = d_best_inter; a.best_score = d_best_score;
    a.coo_row = d_coo_row; a.coo_col = d_coo_col; a.coo_inter = d_coo_inter; a.coo_capacity = coo_capacity;
    a.coo_count = (unsigned long long *)d_coo_count;
    if (d_imat && d_grp_imat_off && imat_ints > 0) {               // dense rows: zeros first, the kernel patches cells
        const cudaError_t e = cudaMemsetAsync(d_imat, 0, (size_t)imat_ints * 4, as_stream(stream));
        if (e != cudaSuccess) { ampis_set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return AMPIS_ECUDA; }
    }
    if (mode == AMPIS_MODE_IOU)
        intersect_rows_grid_kernel<AMPIS_MODE_IOU><<<n_blocks, GR_THREADS, 0, as_stream(stream)>>>(a);
    else
        intersect_rows_grid_kernel<AMPIS_MODE_SAT><<<n_blocks, GR_THREADS, 0, as_stream(stream)>>>(a);
    AMPIS_CHECK_LAUNCH("intersect_rows_grid_kernel");
    return AMPIS_OK;
}
