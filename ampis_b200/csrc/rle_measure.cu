// Per-mask measurements straight from the run counts (one warp per mask), and the
// exclusive scan that turns region sizes into arena offsets.
//
// Replaces pycocotools rleArea (structures.py:568,571; analyze.py:320-321; powder.py:264)
// and yields the tight bounding box extract_boxes (data_utils.py:229-239) would read off the
// decoded mask -- without decoding: a 1-run [s,e) in column-major order lies in column s/h
// if it does not wrap; if it wraps it touches rows 0 and h-1.
#include "common.cuh"
#include "rle_measure.cuh"

__global__ void __launch_bounds__(256)
rle_measure_kernel(const u32 *__restrict__ cnt, const i64 *__restrict__ cnt_off,
                   const int *__restrict__ cnt_len, const u32 *__restrict__ hh, const u32 *__restrict__ ww,
                   int n, int layout, u32 *__restrict__ cum, u32 *__restrict__ area, int *__restrict__ bbox,
                   u32 *__restrict__ span, u32 *__restrict__ reg, i64 *__restrict__ reg_chunks,
                   int *__restrict__ status)
{
    const int i = (int)((blockIdx.x * (u64)blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const i64 base = cnt_off[i];
    const u32 H = hh[i];
    const u64 HW = (u64)H * ww[i];
    const MaskMeasure ms = warp_measure(cnt + base, cnt_len[i], H, HW, nullptr, 0, cum + base);
    if (lane_id() == 0) {
        uint2 sp, rg;
        store_measure(ms, H, HW, layout, i, area, bbox, span, reg, status, &sp, &rg);
        reg_chunks[i] = (i64)(rg.y - rg.x);
    }
}

extern "C" int ampis_rle_measure(const uint32_t *d_cnt, const int64_t *d_cnt_off, const int32_t *d_cnt_len,
                                 const uint32_t *d_h, const uint32_t *d_w, int32_t n, int32_t layout,
                                 uint32_t *d_cum, uint32_t *d_area, int32_t *d_bbox, uint32_t *d_span,
                                 uint32_t *d_reg, int64_t *d_reg_chunks, int32_t *d_status, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    AMPIS_REQUIRE(layout == AMPIS_LAYOUT_SPAN || layout == AMPIS_LAYOUT_FULL || layout == AMPIS_LAYOUT_CROP,
                  "bad layout");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_cnt && d_cnt_off && d_cnt_len && d_h && d_w && d_cum && d_area && d_bbox && d_span &&
                      d_reg && d_reg_chunks && d_status, "null pointer");
    const int warps_per_block = 8;
    const unsigned blocks = (unsigned)((n + warps_per_block - 1) / warps_per_block);
    rle_measure_kernel<<<blocks, warps_per_block * 32, 0, as_stream(stream)>>>(
        d_cnt, d_cnt_off, d_cnt_len, d_h, d_w, n, layout, d_cum, d_area, d_bbox, d_span, d_reg, d_reg_chunks,
        d_status);
    AMPIS_CHECK_LAUNCH("rle_measure_kernel");
    return AMPIS_OK;
}

// ---- exclusive scan of int64 (three passes, all coalesced) -------------------------------------

#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ i64 block_exclusive_scan(i64 v, i64 *total, i64 *smem /*[32]*/)
{
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    i64 incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        i64 t = __shfl_up_sync(0xffffffffu, incl, d);
        if ((int)lane >= d) incl += t;
    }
    if (lane == 31) smem[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        i64 w = lane < (blockDim.x >> 5) ? smem[lane] : 0;
        i64 wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            i64 t = __shfl_up_sync(0xffffffffu, wi, d);
            if ((int)lane >= d) wi += t;
        }
        smem[lane] = wi - w;   // exclusive warp offsets
        if (lane == 31) smem[32] = wi;
    }
    __syncthreads();
    const i64 off = smem[wid];
    *total = smem[32];
    __syncthreads();
    return off + incl - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_sums_kernel(const i64 *__restrict__ in, i64 n, i64 *__restrict__ tile_sums)
{
    __shared__ i64 sm[33];
    const i64 base = (i64)blockIdx.x * SCAN_TILE;
    i64 s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        i64 idx = base + k * SCAN_THREADS + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    i64 total;
    block_exclusive_scan(s, &total, sm);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tile_offsets_kernel(i64 *__restrict__ tile_sums, i64 n_tiles)
{
    __shared__ i64 sm[33];
    i64 carry = 0;
    for (i64 base = 0; base < n_tiles; base += SCAN_THREADS) {
        i64 idx = base + threadIdx.x;
        i64 v = idx < n_tiles ? tile_sums[idx] : 0;
        i64 total;
        i64 ex = block_exclusive_scan(v, &total, sm);
        if (idx < n_tiles) tile_sums[idx] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sums[n_tiles] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_apply_kernel(const i64 *__restrict__ in, i64 n, const i64 *__restrict__ tile_off, i64 n_tiles,
                  i64 *__restrict__ out)
{
    __shared__ i64 sm[33];
    // thread t owns SCAN_ITEMS consecutive items so the in-thread order is the array order
    const i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_ITEMS;
    i64 v[SCAN_ITEMS];
    i64 s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    i64 total;
    i64 ex = block_exclusive_scan(s, &total, sm) + tile_off[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = tile_off[n_tiles];
}

extern "C" size_t ampis_scan_tmp_bytes(int64_t n)
{
    i64 tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return (size_t)(tiles + 1) * sizeof(i64);
}

extern "C" int ampis_exclusive_scan_i64(const int64_t *d_in, int64_t *d_out, int64_t n, void *d_tmp,
                                        size_t tmp_bytes, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    AMPIS_REQUIRE(d_out, "null output");
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(d_out, 0, sizeof(i64), as_stream(stream));
        if (e != cudaSuccess) { ampis_set_error("scan memset: %s", cudaGetErrorString(e)); return AMPIS_ECUDA; }
        return AMPIS_OK;
    }
    AMPIS_REQUIRE(d_in && d_tmp, "null pointer");
    if (tmp_bytes < ampis_scan_tmp_bytes(n)) { ampis_set_error("scan: tmp too small"); return AMPIS_ENOSPC; }
    const i64 tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    i64 *tile_sums = (i64 *)d_tmp;
    scan_tile_sums_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, as_stream(stream)>>>(d_in, n, tile_sums);
    AMPIS_CHECK_LAUNCH("scan_tile_sums_kernel");
    scan_tile_offsets_kernel<<<1, SCAN_THREADS, 0, as_stream(stream)>>>(tile_sums, tiles);
    AMPIS_CHECK_LAUNCH("scan_tile_offsets_kernel");
    scan_apply_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, as_stream(stream)>>>(d_in, n, tile_sums, tiles, d_out);
    AMPIS_CHECK_LAUNCH("scan_apply_kernel");
    return AMPIS_OK;
}
