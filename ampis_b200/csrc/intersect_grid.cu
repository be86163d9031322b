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
//   grid_setup_kernel   CTA per image: cell shift from the column boxes, clears the cell counters
//   grid_count_kernel   thread per column mask: += 1 in every cell its box touches
//   (exclusive scan of the counters by scan_*_kernel: cell -> first entry)
//   grid_fill_kernel    thread per column mask: writes its index into every cell it touches
//   intersect_rows_grid_kernel   warp per row: the cells of the row's box are read by one lane each,
//       their entry lists are flattened by a warp scan, lanes test one entry each (a pair is taken
//       only in the cell holding the top-left corner of the two boxes' overlap, so it is seen once),
//       candidates are intersected four at a time, eight lanes per candidate, as in intersect_crop.cu.
// Entries inside a cell are in no particular order: every arg-max tie is broken explicitly on the
// column index, so the result does not depend on it.
// Optional sparse output: (row, column, intersection) triplets of the non-zero intersections appended
// through an atomic cursor -- the "bbox-pruned sparse IoU" form for images whose dense G x P matrix
// (100 MB at 5,000 x 5,000) is not wanted.
#include "common.cuh"
#include "crop_common.cuh"

#define GR_N 32                 // cells per axis
#define GR_CELLS (GR_N * GR_N)
#define GR_ROWS 8
#define GR_LIST 64
#define GR_BIG 256              // overlap words from which a candidate gets the whole warp

__device__ __forceinline__ int bits_of(u32 x) { return 32 - __clz(x); }
__device__ __forceinline__ int cell_of(int v, int shift) { return min(v >> shift, GR_N - 1); }

__global__ void __launch_bounds__(256)
grid_setup_kernel(const int4 *bbox, const int *grp_col_begin, const int *grp_col_count, int *grp_shift,
                  i64 *cell_count, u32 *cell_fill)
{
    __shared__ unsigned long long s_sum;
    __shared__ u32 s_n, s_ext;
    const int g = blockIdx.x;
    if (threadIdx.x == 0) { s_sum = 0; s_n = 0; s_ext = 0; }
    __syncthreads();
    const int cb = grp_col_begin[g], P = grp_col_count[g];
    u32 sum = 0, n = 0, ext = 0;
    for (int k = threadIdx.x; k < P; k += blockDim.x) {
        const int4 b = bbox[cb + k];
        if (b.z < b.x) continue;
        sum += (u32)max(b.z - b.x, b.w - b.y) + 1u;
        ext = max(ext, (u32)max(b.z, b.w));
        n++;
    }
    sum = warp_sum(sum); n = warp_sum(n); ext = warp_max(ext);
    if (lane_id() == 0) { atomicAdd(&s_sum, (unsigned long long)sum); atomicAdd(&s_n, n); atomicMax(&s_ext, ext); }
    for (int c = threadIdx.x; c < GR_CELLS; c += blockDim.x) {
        cell_count[(i64)g * GR_CELLS + c] = 0;
        cell_fill[(i64)g * GR_CELLS + c] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const u32 mean = s_n ? (u32)((s_sum + s_n - 1) / s_n) : 1u;
        const int by_frame = max(0, bits_of(s_ext) - 5);               // extent >> shift < 32
        const int by_size = bits_of(max(mean, 1u) - 1u);                // 2^shift >= mean box side
        grp_shift[g] = min(max(by_frame, by_size), 30);
    }
}

template <bool FILL>
__global__ void __launch_bounds__(256)
grid_bin_kernel(const int4 *bbox, const int *grp_col_begin, const int *grp_col_count, const int *grp_shift,
                i64 *cell_count, const i64 *cell_off, u32 *cell_fill, int *entries, i64 capacity)
{
    const int g = blockIdx.y;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= grp_col_count[g]) return;
    const int4 b = bbox[grp_col_begin[g] + k];
    if (b.z < b.x) return;
    const int s = grp_shift[g];
    const int cx0 = cell_of(b.x, s), cx1 = cell_of(b.z, s), cy0 = cell_of(b.y, s), cy1 = cell_of(b.w, s);
    for (int cy = cy0; cy <= cy1; cy++)
        for (int cx = cx0; cx <= cx1; cx++) {
            const i64 cell = (i64)g * GR_CELLS + cy * GR_N + cx;
            if (FILL) {
                const i64 pos = cell_off[cell] + (i64)atomicAdd(cell_fill + cell, 1u);
                if (pos < capacity) entries[pos] = k;
            } else {
                atomicAdd(reinterpret_cast<unsigned long long *>(cell_count + cell), 1ull);
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

template <int MODE>
__global__ void __launch_bounds__(GR_ROWS * 32, 4)
intersect_rows_grid_kernel(const GridRowArgs p)
{
    __shared__ int s_cand[GR_ROWS][GR_LIST];
    __shared__ int s_pre[GR_ROWS][32];          // exclusive prefix of the entry counts of the warp's 32 cells
    __shared__ int s_base[GR_ROWS][32];         // first entry of the cell (relative to the image) - prefix

    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const int g = p.blk_grp[blockIdx.x];
    const int r = p.blk_row0[blockIdx.x] + (int)wid;
    if (r >= p.grp_row_begin[g] + p.grp_row_count[g]) return;      // no CTA-wide barrier below
    const int cb = p.grp_col_begin[g];
    const int P = p.grp_col_count[g];
    const i64 imat_off = (p.imat && p.grp_imat_off) ? p.grp_imat_off[g] : -1;
    int *irow = imat_off >= 0 ? p.imat + imat_off + (i64)(r - p.grp_row_begin[g]) * P : nullptr;

    const int rm = p.row_mask[r];
    const int4 rb = p.bbox[rm];
    const u32 ra = p.area[rm];
    const u32 *A = p.words + p.bits_off[rm] * 4;
    double best_s = 0.0;
    u32 best_i = 0;
    int best_c = MODE == AMPIS_MODE_IOU ? -1 : (P > 0 ? 0 : -1);
    int *list = s_cand[wid];

    if (irow) {                                                    // dense row: zeros now, candidates patch later
        for (int k = (int)lane; k < P; k += 32) irow[k] = 0;
        __syncwarp();
    }

    auto flush = [&](int n) {
        for (int j0 = 0; j0 < n; j0 += 4) {
            const int j = j0 + (int)(lane >> 3);
            const bool have = j < n;
            const int k = have ? list[j] : 0;
            Overlap o;
            o.total = 0;
            if (have) o = overlap_of(A, rb, p.words + p.bits_off[cb + k] * 4, p.bbox[cb + k]);
            u32 inter = 0;
            if (__any_sync(0xffffffffu, have && o.total >= GR_BIG)) {
                for (int q = 0; q < 4 && j0 + q < n; q++) {
                    const int kq = list[j0 + q];
                    const Overlap oq = overlap_of(A, rb, p.words + p.bits_off[cb + kq] * 4, p.bbox[cb + kq]);
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
            if (have && (lane & 7u) == 0 && inter) {
                if (irow) irow[k] = (int)inter;
                if (p.coo_count) {
                    const unsigned long long pos = atomicAdd(p.coo_count, 1ull);
                    if ((i64)pos < p.coo_capacity) {
                        p.coo_row[pos] = r; p.coo_col[pos] = k; p.coo_inter[pos] = inter;
                    }
                }
                if (MODE == AMPIS_MODE_IOU) {
                    const double s = (double)inter / (double)(ra + p.area[cb + k] - inter);
                    if (s > best_s || (s == best_s && (unsigned)k < (unsigned)best_c)) {
                        best_s = s; best_i = inter; best_c = k;
                    }
                } else {
                    if (inter > best_i || (inter == best_i && (unsigned)k < (unsigned)best_c)) {
                        best_i = inter; best_c = k;
                    }
                }
            }
        }
    };

    if (ra != 0 && P > 0) {
        const int s = p.grp_shift[g];
        const i64 *off = p.cell_off + (i64)g * GR_CELLS;
        const i64 gbase = off[0];
        const int rcx0 = cell_of(rb.x, s), rcy0 = cell_of(rb.y, s);
        const int ncx = cell_of(rb.z, s) - rcx0 + 1, ncell = ncx * (cell_of(rb.w, s) - rcy0 + 1);
        int n = 0;
        for (int c0 = 0; c0 < ncell; c0 += 32) {
            // one cell per lane: entry range, flattened by an exclusive warp scan of the lengths
            const int ci = c0 + (int)lane;
            int st = 0, len = 0;
            if (ci < ncell) {
                const int cell = (rcy0 + ci / ncx) * GR_N + rcx0 + ci % ncx;
                st = (int)(off[cell] - gbase);
                len = (int)(off[cell + 1] - gbase) - st;
            }
            int incl = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane >= d) incl += t;
            }
            const int T = __shfl_sync(0xffffffffu, incl, 31);
            __syncwarp();                                          // previous round's readers are done
            s_pre[wid][lane] = incl - len;
            s_base[wid][lane] = st - (incl - len);
            __syncwarp();
            for (int t0 = 0; t0 < T; t0 += 32) {
                const int t = t0 + (int)lane;
                bool cand = false;
                int k = 0;
                if (t < T) {
                    int j = 0;                                     // last cell with prefix <= t (empty cells share a prefix)
#pragma unroll
                    for (int d = 16; d; d >>= 1)
                        if (s_pre[wid][j + d] <= t) j += d;
                    const i64 e = gbase + s_base[wid][j] + t;
                    if (e < p.capacity) {
                        k = p.entries[e];
                        const int4 b = p.bbox[cb + k];
                        const int cj = c0 + j;
                        cand = b.x <= rb.z && b.z >= rb.x && b.y <= rb.w && b.w >= rb.y &&
                               max(rcx0, cell_of(b.x, s)) == rcx0 + cj % ncx &&
                               max(rcy0, cell_of(b.y, s)) == rcy0 + cj / ncx;
                    }
                }
                const u32 bal = __ballot_sync(0xffffffffu, cand);
                if (cand) list[n + __popc(bal & ((1u << lane) - 1u))] = k;
                n += __popc(bal);
                if (n > GR_LIST - 32) {
                    __syncwarp();
                    flush(n);
                    n = 0;
                    __syncwarp();
                }
            }
        }
        __syncwarp();
        flush(n);
    }
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

extern "C" int ampis_grid_cells(void) { return GR_CELLS; }

extern "C" int ampis_grid_count(const int32_t *d_bbox, const int32_t *d_grp_col_begin,
                                const int32_t *d_grp_col_count, int32_t n_groups, int32_t max_cols,
                                int32_t *d_grp_shift, int64_t *d_cell_count, uint32_t *d_cell_fill, void *stream)
{
    AMPIS_REQUIRE(n_groups >= 0 && max_cols >= 0, "negative size");
    if (n_groups == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bbox && d_grp_col_begin && d_grp_col_count && d_grp_shift && d_cell_count && d_cell_fill,
                  "null pointer");
    grid_setup_kernel<<<n_groups, 256, 0, as_stream(stream)>>>((const int4 *)d_bbox, d_grp_col_begin,
                                                               d_grp_col_count, d_grp_shift, d_cell_count,
                                                               d_cell_fill);
    AMPIS_CHECK_LAUNCH("grid_setup_kernel");
    if (max_cols == 0) return AMPIS_OK;
    for (int32_t g0 = 0; g0 < n_groups; g0 += 65535) {             // gridDim.y limit
        const dim3 grid((max_cols + 255) / 256, min(n_groups - g0, 65535));
        grid_bin_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(
            (const int4 *)d_bbox, d_grp_col_begin + g0, d_grp_col_count + g0, d_grp_shift + g0,
            d_cell_count + (i64)g0 * GR_CELLS, nullptr, nullptr, nullptr, 0);
        AMPIS_CHECK_LAUNCH("grid_bin_kernel<count>");
    }
    return AMPIS_OK;
}

extern "C" int ampis_grid_fill(const int32_t *d_bbox, const int32_t *d_grp_col_begin,
                               const int32_t *d_grp_col_count, int32_t n_groups, int32_t max_cols,
                               const int32_t *d_grp_shift, const int64_t *d_cell_off, uint32_t *d_cell_fill,
                               int32_t *d_entries, int64_t capacity, void *stream)
{
    AMPIS_REQUIRE(n_groups >= 0 && max_cols >= 0 && capacity >= 0, "negative size");
    if (n_groups == 0 || max_cols == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bbox && d_grp_col_begin && d_grp_col_count && d_grp_shift && d_cell_off && d_cell_fill &&
                      d_entries, "null pointer");
    for (int32_t g0 = 0; g0 < n_groups; g0 += 65535) {
        const dim3 grid((max_cols + 255) / 256, min(n_groups - g0, 65535));
        // cell_off holds absolute entry positions, so only the per-group arrays are shifted
        grid_bin_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(
            (const int4 *)d_bbox, d_grp_col_begin + g0, d_grp_col_count + g0, d_grp_shift + g0, nullptr,
            d_cell_off + (i64)g0 * GR_CELLS, d_cell_fill + (i64)g0 * GR_CELLS, d_entries, capacity);
        AMPIS_CHECK_LAUNCH("grid_bin_kernel<fill>");
    }
    return AMPIS_OK;
}

extern "C" int ampis_intersect_rows_grid(const void *d_bits, const int64_t *d_bits_off, const int32_t *d_bbox,
                                         const uint32_t *d_area, const int32_t *d_row_mask,
                                         const int32_t *d_blk_grp, const int32_t *d_blk_row0, int32_t n_blocks,
                                         const int32_t *d_grp_row_begin, const int32_t *d_grp_row_count,
                                         const int32_t *d_grp_col_begin, const int32_t *d_grp_col_count,
                                         const int32_t *d_grp_shift, const int64_t *d_cell_off,
                                         const int32_t *d_entries, int64_t capacity,
                                         const int64_t *d_grp_imat_off, int32_t mode, int32_t *d_imat,
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
    AMPIS_REQUIRE(d_entries || capacity == 0, "entries missing");
    AMPIS_REQUIRE(!d_coo_count || (d_coo_row && d_coo_col && d_coo_inter && coo_capacity >= 0) || coo_capacity == 0,
                  "sparse output arrays missing");
    GridRowArgs a;
    a.words = (const u32 *)d_bits; a.bits_off = d_bits_off; a.bbox = (const int4 *)d_bbox; a.area = d_area;
    a.row_mask = d_row_mask; a.blk_grp = d_blk_grp; a.blk_row0 = d_blk_row0;
    a.grp_row_begin = d_grp_row_begin; a.grp_row_count = d_grp_row_count;
    a.grp_col_begin = d_grp_col_begin; a.grp_col_count = d_grp_col_count;
    a.grp_shift = d_grp_shift; a.cell_off = d_cell_off; a.entries = d_entries; a.capacity = capacity;
    a.grp_imat_off = d_grp_imat_off; a.imat = d_imat;
    a.best_col = d_best_col; a.best_inter = d_best_inter; a.best_score = d_best_score;
    a.coo_row = d_coo_row; a.coo_col = d_coo_col; a.coo_inter = d_coo_inter; a.coo_capacity = coo_capacity;
    a.coo_count = (unsigned long long *)d_coo_count;
    if (mode == AMPIS_MODE_IOU)
        intersect_rows_grid_kernel<AMPIS_MODE_IOU><<<n_blocks, GR_ROWS * 32, 0, as_stream(stream)>>>(a);
    else
        intersect_rows_grid_kernel<AMPIS_MODE_SAT><<<n_blocks, GR_ROWS * 32, 0, as_stream(stream)>>>(a);
    AMPIS_CHECK_LAUNCH("intersect_rows_grid_kernel");
    return AMPIS_OK;
}
