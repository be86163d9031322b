// Region properties of instance masks (InstanceSet.compute_rprops, structures.py:474-514: the
// reference decodes every mask to a full int64 frame and calls skimage.measure.regionprops_table on
// it, "~30 s" for 2,300 masks).  Nothing is decoded here:
//   * raw moments up to order 2 come straight from the run table (a 1-run is a vertical segment
//     x, [a,b): its contributions are closed-form integer sums) -> centroid, inertia tensor,
//     major/minor axis length, orientation, eccentricity on the host from EXACT integer sums;
//   * the perimeter (skimage.measure.perimeter, 4-neighbourhood: border = mask minus its erosion,
//     each border pixel weighted by how many border pixels touch it by edge / by corner) is
//     bit-sliced arithmetic on the CROP-layout window words;
//   * the convex area (skimage convex_hull_image: hull of the four edge midpoints of every pixel,
//     then a crossing-number point-in-polygon test of the pixel centres) is exact integer geometry:
//     only the top and bottom pixel of every column can touch the hull, two monotone chains in
//     doubled coordinates, and per column the number of pixel centres in [top, bottom).
#include "common.cuh"

// ---- moments from runs: warp per mask ---------------------------------------------------------
// out[i] = { N, sum x, sum y, sum x^2, sum y^2, sum x*y }  (x = column, y = row), exact u64
__global__ void __launch_bounds__(256)
rle_moments_kernel(const u32 *__restrict__ cum, const i64 *__restrict__ cnt_off,
                   const int *__restrict__ cnt_len, const u32 *__restrict__ hh, int n,
                   unsigned long long *__restrict__ out)
{
    const int i = (int)((blockIdx.x * (u32)blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const u32 lane = lane_id();
    const u32 *C = cum + cnt_off[i];
    const int m = cnt_len[i];
    const u64 H = hh[i];
    unsigned long long s[6] = {0, 0, 0, 0, 0, 0};
    for (int r = 1 + 2 * (int)lane; r < m; r += 64) {            // odd runs are the 1-runs
        u64 p = C[r - 1];
        const u64 e = C[r];
        while (p < e) {
            const u64 x = p / H, cs = x * H;
            const long long a = (long long)(p - cs), b = (long long)(min(e, cs + H) - cs);   // rows [a,b)
            const unsigned long long cnt = (unsigned long long)(b - a);
            const unsigned long long sy = (unsigned long long)((a + b - 1) * (b - a) / 2);
            const long long fb = (b - 1) * b * (2 * b - 1) / 6, fa = (a - 1) * a * (2 * a - 1) / 6;
            s[0] += cnt;
            s[1] += cnt * x;
            s[2] += sy;
            s[3] += cnt * x * x;
            s[4] += (unsigned long long)(fb - fa);
            s[5] += sy * x;
            p = cs + H;
        }
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
#pragma unroll
        for (int d = 16; d; d >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], d);
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 6; k++) out[6 * (i64)i + k] = s[k];
    }
}

extern "C" int ampis_rle_moments(const uint32_t *d_cum, const int64_t *d_cnt_off, const int32_t *d_cnt_len,
                                 const uint32_t *d_h, int32_t n, uint64_t *d_out, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_cum && d_cnt_off && d_cnt_len && d_h && d_out, "null pointer");
    rle_moments_kernel<<<(unsigned)(((i64)n * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
        d_cum, d_cnt_off, d_cnt_len, d_h, n, (unsigned long long *)d_out);
    AMPIS_CHECK_LAUNCH("rle_moments_kernel");
    return AMPIS_OK;
}

// ---- CROP window access -------------------------------------------------------------------------
struct Win {
    const u32 *W;
    int x0, x1;        // columns (inclusive)
    u32 wy0, nwy;      // absolute 32-row bands
    __device__ __forceinline__ u32 word(int x, int wy) const     // wy relative to wy0; zero outside
    {
        if (x < x0 || x > x1 || wy < 0 || wy >= (int)nwy) return 0u;
        return W[(u32)(x - x0) * nwy + (u32)wy];
    }
};

__device__ __forceinline__ Win win_of(const u32 *bits, const i64 *bits_off, const int4 *bbox, int i)
{
    const int4 bb = bbox[i];
    Win w;
    w.W = bits + bits_off[i] * 4;
    w.x0 = bb.x; w.x1 = bb.z;
    w.wy0 = (u32)bb.y >> 5;
    w.nwy = bb.z < bb.x ? 0u : ((u32)bb.w >> 5) - w.wy0 + 1u;
    return w;
}

// neighbours of every pixel of word (x, wy): the pixel one row down (y+1) / up (y-1)
__device__ __forceinline__ u32 shift_next_row(const Win &w, int x, int wy)   // bit j <- pixel y+1
{
    return (w.word(x, wy) >> 1) | (w.word(x, wy + 1) << 31);
}
__device__ __forceinline__ u32 shift_prev_row(const Win &w, int x, int wy)   // bit j <- pixel y-1
{
    return (w.word(x, wy) << 1) | (w.word(x, wy - 1) >> 31);
}

// pass 1: border = mask & ~erosion by the 4-neighbourhood cross (outside the box = background)
__global__ void __launch_bounds__(256)
crop_border_kernel(const u32 *__restrict__ bits, const i64 *__restrict__ bits_off, const int4 *__restrict__ bbox,
                   int n, u32 *__restrict__ border)
{
    const int i = (int)((blockIdx.x * (u32)blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const Win w = win_of(bits, bits_off, bbox, i);
    if (!w.nwy) return;
    u32 *B = border + bits_off[i] * 4;
    const u32 total = (u32)(w.x1 - w.x0 + 1) * w.nwy;
    for (u32 idx = lane_id(); idx < total; idx += 32) {
        const int x = w.x0 + (int)(idx / w.nwy), wy = (int)(idx % w.nwy);
        const u32 cur = w.word(x, wy);
        const u32 er = cur & shift_next_row(w, x, wy) & shift_prev_row(w, x, wy) & w.word(x - 1, wy) & w.word(x + 1, wy);
        B[idx] = cur & ~er;
    }
}

// pass 2: every border pixel gets the code 1 + 2*(edge-adjacent border pixels) + 10*(corner-adjacent
// border pixels) -- the value skimage's 3x3 convolution [[10,2,10],[2,1,2],[10,2,10]] yields -- and
// the ten codes that carry a weight are counted: order 5,7,15,17,25,27 (weight 1), 21,33 (sqrt 2),
// 13,23 ((1+sqrt 2)/2).
__constant__ int c_code_e[10] = {2, 3, 2, 3, 2, 3, 0, 1, 1, 1};
__constant__ int c_code_d[10] = {0, 0, 1, 1, 2, 2, 2, 3, 1, 2};

__device__ __forceinline__ void count4(u32 a, u32 b, u32 c, u32 d, u32 &p0, u32 &p1, u32 &p2)
{
    const u32 s1 = a ^ b, c1 = a & b, s2 = c ^ d, c2 = c & d, t = s1 & s2;
    p0 = s1 ^ s2; p1 = c1 ^ c2 ^ t; p2 = c1 & c2;
}
__device__ __forceinline__ u32 equals_k(u32 p0, u32 p1, u32 p2, int k)
{
    return ((k & 1) ? p0 : ~p0) & ((k & 2) ? p1 : ~p1) & ((k & 4) ? p2 : ~p2);
}

__global__ void __launch_bounds__(256)
crop_perimeter_kernel(const u32 *__restrict__ border, const i64 *__restrict__ bits_off,
                      const int4 *__restrict__ bbox, int n, u32 *__restrict__ hist10)
{
    const int i = (int)((blockIdx.x * (u32)blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const Win w = win_of(border, bits_off, bbox, i);
    u32 h[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const u32 total = w.nwy ? (u32)(w.x1 - w.x0 + 1) * w.nwy : 0u;
    for (u32 idx = lane_id(); idx < total; idx += 32) {
        const int x = w.x0 + (int)(idx / w.nwy), wy = (int)(idx % w.nwy);
        const u32 bc = w.word(x, wy);
        if (!bc) continue;
        u32 e0, e1, e2, d0, d1, d2;
        count4(shift_next_row(w, x, wy), shift_prev_row(w, x, wy), w.word(x - 1, wy), w.word(x + 1, wy), e0, e1, e2);
        count4(shift_next_row(w, x - 1, wy), shift_prev_row(w, x - 1, wy), shift_next_row(w, x + 1, wy),
               shift_prev_row(w, x + 1, wy), d0, d1, d2);
#pragma unroll
        for (int k = 0; k < 10; k++)
            h[k] += __popc(bc & equals_k(e0, e1, e2, c_code_e[k]) & equals_k(d0, d1, d2, c_code_d[k]));
    }
#pragma unroll
    for (int k = 0; k < 10; k++) {
        const u32 v = warp_sum(h[k]);
        if (lane_id() == 0) hist10[10 * (i64)i + k] = v;
    }
}

// ---- convex area: thread per mask -----------------------------------------------------------------
// Doubled coordinates X = 2*column, Y = 2*row.  Candidates of the lower chain (smallest row per X) and
// of the upper chain (largest row per X) come from the top / bottom pixel of every column:
//   even X = 2x   : top(x)*2 - 1            / bottom(x)*2 + 1
//   odd  X = 2x+1 : min top of x, x+1 (*2)  / max bottom of x, x+1 (*2)
__device__ __forceinline__ i64 ceil_div(i64 a, i64 b)      // b > 0
{
    return a >= 0 ? (a + b - 1) / b : -((-a) / b);
}

__global__ void __launch_bounds__(128)
crop_convex_area_kernel(const u32 *__restrict__ bits, const i64 *__restrict__ bits_off,
                        const int4 *__restrict__ bbox, int n, const i64 *__restrict__ scr_off,
                        int *__restrict__ scratch, unsigned long long *__restrict__ convex_area)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Win w = win_of(bits, bits_off, bbox, i);
    if (!w.nwy) { convex_area[i] = 0; return; }
    const int bw = w.x1 - w.x0 + 1;
    // scratch per mask: top[bw], bot[bw], then two chains of (X, Y) with at most 2*bw + 1 points each
    int *top = scratch + scr_off[i], *bot = top + bw;
    int *LX = bot + bw, *LY = LX + (2 * bw + 1), *UX = LY + (2 * bw + 1), *UY = UX + (2 * bw + 1);
    for (int c = 0; c < bw; c++) {
        int t = -1, b = -1;
        for (u32 k = 0; k < w.nwy; k++) {
            const u32 v = w.W[(u32)c * w.nwy + k];
            if (v) { t = (int)((w.wy0 + k) * 32u) + __ffs(v) - 1; break; }
        }
        for (int k = (int)w.nwy - 1; k >= 0 && t >= 0; k--) {
            const u32 v = w.W[(u32)c * w.nwy + (u32)k];
            if (v) { b = (int)((w.wy0 + (u32)k) * 32u) + 31 - __clz(v); break; }
        }
        top[c] = t; bot[c] = b;
    }
    int nl = 0, nu = 0;
    for (int X = -1; X <= 2 * bw - 1; X++) {               // X relative to 2*x0
        int yl, yu;
        if (X & 1) {                                        // between columns c = (X-1)/2 and c+1
            const int c = (X - 1) >> 1;                     // floor for X = -1 gives -1
            const int t0 = (c >= 0 && top[c] >= 0) ? top[c] : 0x3fffffff;
            const int t1 = (c + 1 < bw && top[c + 1] >= 0) ? top[c + 1] : 0x3fffffff;
            const int b0 = (c >= 0 && top[c] >= 0) ? bot[c] : -1;
            const int b1 = (c + 1 < bw && top[c + 1] >= 0) ? bot[c + 1] : -1;
            if (min(t0, t1) == 0x3fffffff) continue;
            yl = 2 * min(t0, t1); yu = 2 * max(b0, b1);
        } else {
            const int c = X >> 1;
            if (top[c] < 0) continue;
            yl = 2 * top[c] - 1; yu = 2 * bot[c] + 1;
        }
        // lower chain: slopes must increase; upper chain: slopes must decrease
        while (nl >= 2 && (i64)(LX[nl - 1] - LX[nl - 2]) * (yl - LY[nl - 2]) -
                              (i64)(LY[nl - 1] - LY[nl - 2]) * (X - LX[nl - 2]) <= 0) nl--;
        LX[nl] = X; LY[nl] = yl; nl++;
        while (nu >= 2 && (i64)(UX[nu - 1] - UX[nu - 2]) * (yu - UY[nu - 2]) -
                              (i64)(UY[nu - 1] - UY[nu - 2]) * (X - UX[nu - 2]) >= 0) nu--;
        UX[nu] = X; UY[nu] = yu; nu++;
    }
    // pixel centres inside: per column c the rows r with lower(c) <= r < upper(c) (crossing-number rule)
    unsigned long long area = 0;
    int il = 0, iu = 0;
    for (int c = 0; c < bw; c++) {
        const int X = 2 * c;
        while (il + 2 < nl && LX[il + 1] <= X) il++;
        while (iu + 2 < nu && UX[iu + 1] <= X) iu++;
        if (X < LX[0] || X > LX[nl - 1]) continue;          // column outside the hull (cannot happen for c in box)
        i64 lo, hi;
        {
            const i64 D = LX[il + 1] - LX[il];
            const i64 num = (i64)LY[il] * D + (i64)(LY[il + 1] - LY[il]) * (X - LX[il]);    // = Y * D, row = Y / 2
            lo = ceil_div(num, 2 * D);
        }
        {
            const i64 D = UX[iu + 1] - UX[iu];
            const i64 num = (i64)UY[iu] * D + (i64)(UY[iu + 1] - UY[iu]) * (X - UX[iu]);
            hi = ceil_div(num, 2 * D);
        }
        if (hi > lo) area += (unsigned long long)(hi - lo);
    }
    convex_area[i] = area;
}

extern "C" int ampis_crop_perimeter(const void *d_bits, const int64_t *d_bits_off, const int32_t *d_bbox,
                                    int32_t n, void *d_border_scratch, uint32_t *d_hist10, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits && d_bits_off && d_bbox && d_border_scratch && d_hist10, "null pointer");
    const unsigned blocks = (unsigned)(((i64)n * 32 + 255) / 256);
    crop_border_kernel<<<blocks, 256, 0, as_stream(stream)>>>((const u32 *)d_bits, d_bits_off, (const int4 *)d_bbox, n,
                                                              (u32 *)d_border_scratch);
    AMPIS_CHECK_LAUNCH("crop_border_kernel");
    crop_perimeter_kernel<<<blocks, 256, 0, as_stream(stream)>>>((const u32 *)d_border_scratch, d_bits_off,
                                                                 (const int4 *)d_bbox, n, d_hist10);
    AMPIS_CHECK_LAUNCH("crop_perimeter_kernel");
    return AMPIS_OK;
}

extern "C" int ampis_crop_convex_area(const void *d_bits, const int64_t *d_bits_off, const int32_t *d_bbox,
                                      int32_t n, const int64_t *d_scratch_off, int32_t *d_scratch,
                                      uint64_t *d_convex_area, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits && d_bits_off && d_bbox && d_scratch_off && d_scratch && d_convex_area, "null pointer");
    crop_convex_area_kernel<<<(unsigned)((n + 127) / 128), 128, 0, as_stream(stream)>>>(
        (const u32 *)d_bits, d_bits_off, (const int4 *)d_bbox, n, d_scratch_off, d_scratch,
        (unsigned long long *)d_convex_area);
    AMPIS_CHECK_LAUNCH("crop_convex_area_kernel");
    return AMPIS_OK;
}
