// Label / binary annotation images -> per-instance RLE (data_utils.get_ddicts, label_fmt 'binary' and
// 'label', data_utils.py:394-433): the reference calls skimage.measure.label on the binary image
// (full 8-connectivity in 2-D, labels numbered in raster order of each component's first pixel),
// then for every label value u builds the mask `ann == u`, takes its box (extract_boxes) and
// RLE.encode()s it -- one full-frame pass per instance.  Here: union-find labelling, one dense
// relabel written column-major (COCO order), per-label boxes by atomics, and a warp per label that
// walks only the label's box window and emits its run boundaries in order.
#include "common.cuh"

// ---- connected components, 8-connectivity ---------------------------------------------------------
__device__ __forceinline__ int uf_find(const int *L, int i)
{
    int p = ((const volatile int *)L)[i];
    while (p != i) { i = p; p = ((const volatile int *)L)[i]; }
    return i;
}
__device__ __forceinline__ void uf_union(int *L, int a, int b)
{
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }      // link the larger root under the smaller
        const int old = atomicMin(&L[a], b);
        if (old == a) return;
        a = old;
    }
}

__global__ void __launch_bounds__(256)
ccl_init_kernel(const uint8_t *__restrict__ img, i64 n, int *__restrict__ L)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) L[i] = img[i] ? (int)i : -1;
}

__global__ void __launch_bounds__(256)
ccl_merge_kernel(const uint8_t *__restrict__ img, int h, int w, int *L)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (i64)h * w || !img[i]) return;
    const int y = (int)(i / w), x = (int)(i - (i64)y * w);
    if (x > 0 && img[i - 1]) uf_union(L, (int)i, (int)i - 1);
    if (y > 0) {
        if (img[i - w]) uf_union(L, (int)i, (int)(i - w));
        if (x > 0 && img[i - w - 1]) uf_union(L, (int)i, (int)(i - w - 1));
        if (x + 1 < w && img[i - w + 1]) uf_union(L, (int)i, (int)(i - w + 1));
    }
}

// root flags (component representative = its smallest row-major index => raster order of first pixels)
__global__ void __launch_bounds__(256)
ccl_roots_kernel(int *L, i64 n, i64 *__restrict__ is_root)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = -1;
    if (L[i] >= 0) r = uf_find(L, (int)i);
    is_root[i] = (r == (int)i) ? 1 : 0;
}

// dense labels 1..n in raster order of the roots, written TRANSPOSED (column-major, COCO pixel order)
__global__ void __launch_bounds__(256)
ccl_relabel_kernel(const int *__restrict__ L, const i64 *__restrict__ root_rank, int h, int w,
                   int *__restrict__ dense_t)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (i64)h * w) return;
    const int y = (int)(i / w), x = (int)(i - (i64)y * w);
    int v = 0;
    if (L[i] >= 0) v = (int)root_rank[uf_find(L, (int)i)] + 1;
    dense_t[(i64)x * h + y] = v;
}

// ---- arbitrary label values -> dense ids in ascending value order ('label' format) -----------------
__global__ void __launch_bounds__(256)
label_present_kernel(const int *__restrict__ ann, i64 n, i64 *__restrict__ present, int n_values, int *bad)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = ann[i];
    if (v < 0 || v >= n_values) { *bad = 1; return; }
    present[v] = 1;       // benign race: every writer stores 1
}
__global__ void __launch_bounds__(256)
label_dense_kernel(const int *__restrict__ ann, const i64 *__restrict__ rank, int zero_present, int h, int w,
                   int *__restrict__ dense_t)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (i64)h * w) return;
    const int y = (int)(i / w), x = (int)(i - (i64)y * w);
    // np.unique order; the smallest value is skipped when it is 0 (data_utils.py:412-414)
    dense_t[(i64)x * h + y] = (int)rank[ann[i]] + (zero_present ? 0 : 1);
}

// ---- per-label boxes ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
label_bbox_kernel(const int *__restrict__ dense_t, int h, int w, int *__restrict__ bbox /* x0,y0,x1,y1 */)
{
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= (i64)h * w) return;
    const int u = dense_t[k];
    if (u <= 0) return;
    const int x = (int)(k / h), y = (int)(k - (i64)x * h);
    int *b = bbox + 4 * (i64)(u - 1);
    atomicMin(b + 0, x); atomicMin(b + 1, y); atomicMax(b + 2, x); atomicMax(b + 3, y);
}

// ---- per-label run boundaries: warp per label over its box window, column-major order --------------
// EMIT = false: count boundaries (n_runs = boundaries + 1).  EMIT = true: write boundary positions at
// cnt + cnt_off[u], then turn them into run counts in place.
template <bool EMIT>
__global__ void __launch_bounds__(256)
label_rle_kernel(const int *__restrict__ dense_t, int h, int w, int n_labels, const int *__restrict__ bbox,
                 i64 *__restrict__ n_runs, const i64 *__restrict__ cnt_off, u32 *__restrict__ cnt,
                 int *__restrict__ cnt_len)
{
    const int u0 = (int)((blockIdx.x * (u32)blockDim.x + threadIdx.x) >> 5);
    if (u0 >= n_labels) return;
    const u32 lane = lane_id();
    const int u = u0 + 1;
    const int4 bb = reinterpret_cast<const int4 *>(bbox)[u0];
    u32 *pos = EMIT ? cnt + cnt_off[u0] : nullptr;
    u32 nb = 0, prev = 0;
    const u64 hw = (u64)h * w;
    for (int x = bb.x; x <= bb.z; x++) {
        const int *col = dense_t + (i64)x * h;
        for (int y0 = bb.y; y0 <= bb.w; y0 += 32) {
            const int y = y0 + (int)lane;
            const bool v = y <= bb.w && col[y] == u;
            const u32 b = __ballot_sync(0xffffffffu, v);
            const u32 nvalid = (u32)min(32, bb.w - y0 + 1);
            u32 t = b ^ ((b << 1) | prev);
            if (nvalid < 32) t &= (1u << nvalid) - 1u;       // the edge after the last valid row is the gap's
            if (EMIT && ((t >> lane) & 1u)) pos[nb + __popc(t & ((1u << lane) - 1u))] = (u32)((u64)x * h + y);
            nb += __popc(t);
            prev = (b >> (nvalid - 1)) & 1u;
        }
        // pixels between the bottom of this window column and the top of the next are background
        const bool gap = bb.w < h - 1 || bb.y > 0 || x == bb.z;
        if (gap && prev) {
            const u64 k = (u64)x * h + bb.w + 1;
            if (k < hw) {
                if (EMIT && lane == 0) pos[nb] = (u32)k;
                nb++;
            }
            prev = 0;
        }
    }
    if (!EMIT) {
        if (lane == 0) n_runs[u0] = (i64)nb + 1;
        return;
    }
    __syncwarp();
    const u32 m = nb + 1;
    for (i64 hi = (i64)m; hi > 0; hi -= 32) {          // positions -> counts, back to front
        const i64 j = hi - 1 - lane;
        u32 val = 0;
        if (j >= 0) {
            const u32 end = j == (i64)nb ? (u32)hw : pos[j];
            const u32 start = j ? pos[j - 1] : 0u;
            val = end - start;
        }
        __syncwarp();
        if (j >= 0) pos[j] = val;
        __syncwarp();
    }
    if (lane == 0) cnt_len[u0] = (int)m;
}

#define GRID1(n) (unsigned)(((n) + 255) / 256)

extern "C" int ampis_ccl_label(const uint8_t *d_img, int32_t h, int32_t w, int32_t *d_work, int64_t *d_flags,
                               int64_t *d_rank, void *d_scan_tmp, size_t scan_tmp_bytes, int32_t *d_dense_t,
                               void *stream)
{
    AMPIS_REQUIRE(h > 0 && w > 0 && (i64)h * w < (1ll << 31), "bad image size");
    AMPIS_REQUIRE(d_img && d_work && d_flags && d_rank && d_scan_tmp && d_dense_t, "null pointer");
    const i64 n = (i64)h * w;
    cudaStream_t s = as_stream(stream);
    ccl_init_kernel<<<GRID1(n), 256, 0, s>>>(d_img, n, d_work);
    ccl_merge_kernel<<<GRID1(n), 256, 0, s>>>(d_img, h, w, d_work);
    ccl_roots_kernel<<<GRID1(n), 256, 0, s>>>(d_work, n, d_flags);
    AMPIS_CHECK_LAUNCH("ccl kernels");
    int rc = ampis_exclusive_scan_i64(d_flags, d_rank, n, d_scan_tmp, scan_tmp_bytes, stream);
    if (rc) return rc;
    ccl_relabel_kernel<<<GRID1(n), 256, 0, s>>>(d_work, d_rank, h, w, d_dense_t);
    AMPIS_CHECK_LAUNCH("ccl_relabel_kernel");
    return AMPIS_OK;       // number of labels = d_rank[n]
}

extern "C" int ampis_label_values_present(const int32_t *d_ann, int64_t n, int64_t *d_present, int32_t n_values,
                                          int32_t *d_bad, void *stream)
{
    AMPIS_REQUIRE(n >= 0 && n_values > 0, "bad size");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_ann && d_present && d_bad, "null pointer");
    label_present_kernel<<<GRID1(n), 256, 0, as_stream(stream)>>>(d_ann, n, d_present, n_values, d_bad);
    AMPIS_CHECK_LAUNCH("label_present_kernel");
    return AMPIS_OK;
}

extern "C" int ampis_label_dense(const int32_t *d_ann, const int64_t *d_rank, int32_t zero_present, int32_t h,
                                 int32_t w, int32_t *d_dense_t, void *stream)
{
    AMPIS_REQUIRE(h > 0 && w > 0, "bad image size");
    AMPIS_REQUIRE(d_ann && d_rank && d_dense_t, "null pointer");
    label_dense_kernel<<<GRID1((i64)h * w), 256, 0, as_stream(stream)>>>(d_ann, d_rank, zero_present, h, w,
                                                                         d_dense_t);
    AMPIS_CHECK_LAUNCH("label_dense_kernel");
    return AMPIS_OK;
}

extern "C" int ampis_label_bbox(const int32_t *d_dense_t, int32_t h, int32_t w, int32_t n_labels, int32_t *d_bbox,
                                void *stream)
{
    AMPIS_REQUIRE(h > 0 && w > 0 && n_labels >= 0, "bad size");
    if (n_labels == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_dense_t && d_bbox, "null pointer");   // d_bbox pre-filled with (INT_MAX, INT_MAX, -1, -1)
    label_bbox_kernel<<<GRID1((i64)h * w), 256, 0, as_stream(stream)>>>(d_dense_t, h, w, d_bbox);
    AMPIS_CHECK_LAUNCH("label_bbox_kernel");
    return AMPIS_OK;
}

extern "C" int ampis_label_rle_count(const int32_t *d_dense_t, int32_t h, int32_t w, int32_t n_labels,
                                     const int32_t *d_bbox, int64_t *d_n_runs, void *stream)
{
    AMPIS_REQUIRE(h > 0 && w > 0 && n_labels >= 0, "bad size");
    if (n_labels == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_dense_t && d_bbox && d_n_runs, "null pointer");
    label_rle_kernel<false><<<GRID1((i64)n_labels * 32), 256, 0, as_stream(stream)>>>(
        d_dense_t, h, w, n_labels, d_bbox, d_n_runs, nullptr, nullptr, nullptr);
    AMPIS_CHECK_LAUNCH("label_rle_kernel<count>");
    return AMPIS_OK;
}

extern "C" int ampis_label_rle_emit(const int32_t *d_dense_t, int32_t h, int32_t w, int32_t n_labels,
                                    const int32_t *d_bbox, const int64_t *d_cnt_off, uint32_t *d_cnt,
                                    int32_t *d_cnt_len, void *stream)
{
    AMPIS_REQUIRE(h > 0 && w > 0 && n_labels >= 0, "bad size");
    if (n_labels == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_dense_t && d_bbox && d_cnt_off && d_cnt && d_cnt_len, "null pointer");
    label_rle_kernel<true><<<GRID1((i64)n_labels * 32), 256, 0, as_stream(stream)>>>(
        d_dense_t, h, w, n_labels, d_bbox, nullptr, d_cnt_off, d_cnt, d_cnt_len);
    AMPIS_CHECK_LAUNCH("label_rle_kernel<emit>");
    return AMPIS_OK;
}
