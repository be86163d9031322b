// Polygon -> RLE on the GPU, bit-exact with pycocotools rleFrPoly,
// which AMPIS reaches through RLE.frPyObjects for polygon ground truth (structures.py:677).
//
// rleFrPoly up-samples the boundary x5, walks every edge with an integer DDA, records where
// the walk crosses pixel-column centres, sorts those crossing positions (column-major pixel
// index), differences them into run lengths and fuses zero-length runs.  The fusing rule is
// "equal positions cancel in pairs" (a position survives iff its multiplicity is odd) and the
// terminal position h*w always survives once -- see DESIGN.md for the derivation.  That makes
// the algorithm data-parallel: every DDA point is computed independently from (edge, step),
// crossings are gathered in shared memory, bitonic-sorted, parity-filtered and differenced.
//
// All double arithmetic uses explicit round-to-nearest intrinsics so that no FMA contraction
// changes a rounding relative to the CPU build of pycocotools (x86-64, no FMA).
#include "common.cuh"

#define POLY_THREADS 256
#define POLY_MAX_VERTS 4096
#define POLY_MAX_CROSS 8192

__device__ __forceinline__ int scale5(double v)   // (int)(5*v + .5), C truncation
{
    return __double2int_rz(__dadd_rn(__dmul_rn(5.0, v), 0.5));
}

struct Pt { int u, v; };

// d-th DDA point of edge (xs,ys)->(xe,ye)
__device__ __forceinline__ Pt dda_point(int xs, int ys, int xe, int ye, int d)
{
    const int dx = abs(xe - xs), dy = abs(ys - ye);
    const bool flip = (dx >= dy && xs > xe) || (dx < dy && ys > ye);
    if (flip) { int t = xs; xs = xe; xe = t; t = ys; ys = ye; ye = t; }
    Pt p;
    if (dx >= dy) {
        const double s = __ddiv_rn((double)(ye - ys), (double)dx);
        const int t = flip ? dx - d : d;
        p.u = t + xs;
        p.v = __double2int_rz(__dadd_rn(__dadd_rn((double)ys, __dmul_rn(s, (double)t)), 0.5));
    } else {
        const double s = __ddiv_rn((double)(xe - xs), (double)dy);
        const int t = flip ? dy - d : d;
        p.v = t + ys;
        p.u = __double2int_rz(__dadd_rn(__dadd_rn((double)xs, __dmul_rn(s, (double)t)), 0.5));
    }
    return p;
}

__global__ void __launch_bounds__(POLY_THREADS)
poly_to_rle_kernel(const double *__restrict__ xy, const i64 *__restrict__ xy_off, const u32 *__restrict__ hh,
                   const u32 *__restrict__ ww, u32 *__restrict__ cnt, const i64 *__restrict__ cnt_off,
                   int *__restrict__ cnt_len)
{
    extern __shared__ u32 sm[];
    int *eoff = reinterpret_cast<int *>(sm);             // [POLY_MAX_VERTS + 1] first point of edge j
    u32 *pos = sm + POLY_MAX_VERTS + 1;                    // [POLY_MAX_CROSS] crossing positions
    u32 *kept = pos + POLY_MAX_CROSS;                      // [POLY_MAX_CROSS] survivors
    __shared__ int s_n, s_scan[POLY_THREADS / 32 + 1];

    const int i = blockIdx.x;
    const double *P = xy + xy_off[i];
    const int k = (int)((xy_off[i + 1] - xy_off[i]) / 2);
    const u32 h = hh[i], w = ww[i];
    const u32 hw = h * w;
    const i64 cap = cnt_off[i + 1] - cnt_off[i];
    u32 *out = cnt + cnt_off[i];
    const int tid = threadIdx.x;

    if (k < 1 || k > POLY_MAX_VERTS) {
        if (tid == 0) cnt_len[i] = -1;
        return;
    }
    // points per edge, then an exclusive scan into eoff[]
    for (int j = tid; j < k; j += POLY_THREADS) {
        const int jn = j + 1 == k ? 0 : j + 1;
        const int dx = abs(scale5(P[2 * j]) - scale5(P[2 * jn]));
        const int dy = abs(scale5(P[2 * j + 1]) - scale5(P[2 * jn + 1]));
        eoff[j] = max(dx, dy) + 1;
    }
    if (tid == 0) s_n = 0;
    __syncthreads();
    {   // serial-by-tiles block scan (k is small)
        int carry = 0;
        for (int base = 0; base < k; base += POLY_THREADS) {
            const int j = base + tid;
            const int v = j < k ? eoff[j] : 0;
            int incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, incl, d);
                if ((int)lane_id() >= d) incl += t;
            }
            if (lane_id() == 31) s_scan[tid >> 5] = incl;
            __syncthreads();
            int woff = 0;
            for (int q = 0; q < (tid >> 5); q++) woff += s_scan[q];
            int tot = 0;
            for (int q = 0; q < POLY_THREADS / 32; q++) tot += s_scan[q];
            if (j < k) eoff[j] = carry + woff + incl - v;
            carry += tot;
            __syncthreads();
        }
        if (tid == 0) eoff[k] = carry;
    }
    __syncthreads();
    const int M = eoff[k];

    // every consecutive pair of DDA points -> at most one crossing
    for (int q = 1 + tid; q < M; q += POLY_THREADS) {
        Pt a[2];
#pragma unroll
        for (int s = 0; s < 2; s++) {
            const int qq = q - 1 + s;
            int lo = 0, hi = k;   // last edge j with eoff[j] <= qq
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (eoff[mid] <= qq) lo = mid; else hi = mid;
            }
            const int j = lo, jn = j + 1 == k ? 0 : j + 1;
            a[s] = dda_point(scale5(P[2 * j]), scale5(P[2 * j + 1]), scale5(P[2 * jn]), scale5(P[2 * jn + 1]),
                             qq - eoff[j]);
        }
        if (a[1].u == a[0].u) continue;
        double xd = (double)(a[1].u < a[0].u ? a[1].u : a[1].u - 1);
        xd = __dadd_rn(__ddiv_rn(__dadd_rn(xd, 0.5), 5.0), -0.5);
        if (floor(xd) != xd || xd < 0 || xd > (double)w - 1) continue;
        double yd = (double)(a[1].v < a[0].v ? a[1].v : a[0].v);
        yd = __dadd_rn(__ddiv_rn(__dadd_rn(yd, 0.5), 5.0), -0.5);
        if (yd < 0) yd = 0; else if (yd > (double)h) yd = (double)h;
        yd = ceil(yd);
        const u32 v = (u32)(__double2int_rz(xd) * (int)h + __double2int_rz(yd));
        if (v >= hw) continue;   // copies of the terminal position never change the result
        const int slot = atomicAdd(&s_n, 1);
        if (slot < POLY_MAX_CROSS) pos[slot] = v;
    }
    __syncthreads();
    const int n = s_n;
    if (n > POLY_MAX_CROSS) {
        if (tid == 0) cnt_len[i] = -1;
        return;
    }
    // bitonic sort of pos[0..n2), padded with +inf
    int n2 = 1;
    while (n2 < n) n2 <<= 1;
    for (int q = n + tid; q < n2; q += POLY_THREADS) pos[q] = 0xffffffffu;
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int q = tid; q < n2 / 2; q += POLY_THREADS) {
                const int lo = 2 * q - (q & (stride - 1));
                const int hi = lo + stride;
                const bool up = (lo & size) == 0;
                const u32 x = pos[lo], y = pos[hi];
                if ((x > y) == up) { pos[lo] = y; pos[hi] = x; }
            }
            __syncthreads();
        }
    // survivors: last element of each group of equal values, if the group has odd size
    if (tid == 0) s_n = 0;
    __syncthreads();
    int nk_total = 0;
    for (int base = 0; base < n; base += POLY_THREADS) {
        const int q = base + tid;
        bool keep = false;
        if (q < n && (q + 1 == n || pos[q + 1] != pos[q])) {
            int lo = 0, hi = q;   // first index with pos == pos[q]
            const u32 v = pos[q];
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (pos[mid] < v) lo = mid + 1; else hi = mid;
            }
            keep = ((q - lo + 1) & 1) != 0;
        }
        const u32 bal = __ballot_sync(0xffffffffu, keep);
        if (lane_id() == 0) s_scan[tid >> 5] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int qq = 0; qq < POLY_THREADS / 32; qq++) {
            if (qq < (tid >> 5)) woff += s_scan[qq];
            tot += s_scan[qq];
        }
        if (keep) kept[nk_total + woff + __popc(bal & ((1u << lane_id()) - 1u))] = pos[q];
        nk_total += tot;
        __syncthreads();
    }
    const int m = nk_total + 1;
    if (tid == 0) cnt_len[i] = m;
    if ((i64)m > cap) return;
    for (int q = tid; q < m; q += POLY_THREADS) {
        const u32 prev = q == 0 ? 0u : kept[q - 1];
        const u32 cur = q == nk_total ? hw : kept[q];
        out[q] = cur - prev;
    }
}

extern "C" int ampis_poly_to_rle(const double *d_xy, const int64_t *d_xy_off, const uint32_t *d_h,
                                 const uint32_t *d_w, int32_t n, uint32_t *d_cnt, const int64_t *d_cnt_off,
                                 int32_t *d_cnt_len, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_xy && d_xy_off && d_h && d_w && d_cnt && d_cnt_off && d_cnt_len, "null pointer");
    const size_t smem = (size_t)(POLY_MAX_VERTS + 1 + 2 * POLY_MAX_CROSS) * sizeof(u32);
    // per launch: the attribute is per device, and a process may drive more than one
    cudaError_t e = cudaFuncSetAttribute(poly_to_rle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { ampis_set_error("poly smem attr: %s", cudaGetErrorString(e)); return AMPIS_ECUDA; }
    poly_to_rle_kernel<<<n, POLY_THREADS, smem, as_stream(stream)>>>(d_xy, d_xy_off, d_h, d_w, d_cnt, d_cnt_off,
                                                                    d_cnt_len);
    AMPIS_CHECK_LAUNCH("poly_to_rle_kernel");
    return AMPIS_OK;
}

// ---- skimage.draw.polygon2mask (structures._poly2mask, structures.py:693-715) ----------------------
// masks_to_bitmask_array(PolygonMasks) in the reference does NOT use pycocotools' rasteriser but
// skimage's: every pixel (r, c) of the polygon's integer bounding range is tested with the
// crossing-number rule of skimage/measure/_pnpoly.pxd on (x = c, y = r):
//     inside ^= ((yp[i] <= y < yp[j]) or (yp[j] <= y < yp[i])) and
//               x < (xp[j] - xp[i]) * (y - yp[i]) / (yp[j] - yp[i]) + xp[i]
// evaluated in double precision.  The same expression with explicitly rounded operations (no FMA
// contraction) gives the same bits.  CTA per polygon, threads stride over the bounding range.
__global__ void __launch_bounds__(256)
polygon2mask_kernel(const double *__restrict__ xy, const i64 *__restrict__ xy_off, int h, int w,
                    uint8_t *__restrict__ out)
{
    const int i = blockIdx.x;
    const double *P = xy + xy_off[i];
    const int nv = (int)((xy_off[i + 1] - xy_off[i]) / 2);
    if (nv < 1) return;
    __shared__ double s_lim[4];
    if (threadIdx.x == 0) {
        double xmin = P[0], xmax = P[0], ymin = P[1], ymax = P[1];
        for (int k = 1; k < nv; k++) {
            xmin = fmin(xmin, P[2 * k]); xmax = fmax(xmax, P[2 * k]);
            ymin = fmin(ymin, P[2 * k + 1]); ymax = fmax(ymax, P[2 * k + 1]);
        }
        s_lim[0] = xmin; s_lim[1] = xmax; s_lim[2] = ymin; s_lim[3] = ymax;
    }
    __syncthreads();
    // minr = int(max(0, r.min())), maxr = min(h - 1, int(ceil(r.max()))), same for columns
    const long long minr = (long long)fmax(0.0, s_lim[2]), maxr = min((long long)h - 1, (long long)ceil(s_lim[3]));
    const long long minc = (long long)fmax(0.0, s_lim[0]), maxc = min((long long)w - 1, (long long)ceil(s_lim[1]));
    if (maxr < minr || maxc < minc) return;
    const long long nc = maxc - minc + 1, total = (maxr - minr + 1) * nc;
    uint8_t *o = out + (i64)i * h * w;
    for (long long idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const long long r = minr + idx / nc, c = minc + idx % nc;
        const double x = (double)c, y = (double)r;
        bool in = false;
        int j = nv - 1;
        for (int k = 0; k < nv; k++) {
            const double xi = P[2 * k], yi = P[2 * k + 1], xj = P[2 * j], yj = P[2 * j + 1];
            if (((yi <= y) && (y < yj)) || ((yj <= y) && (y < yi))) {
                const double t = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(xj, xi), __dsub_rn(y, yi)), __dsub_rn(yj, yi)), xi);
                if (x < t) in = !in;
            }
            j = k;
        }
        if (in) o[r * w + c] = 1;
    }
}

extern "C" int ampis_polygon2mask(const double *d_xy, const int64_t *d_xy_off, int32_t n, int32_t h, int32_t w,
                                  uint8_t *d_out, void *stream)
{
    AMPIS_REQUIRE(n >= 0 && h > 0 && w > 0, "bad size");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_xy && d_xy_off && d_out, "null pointer");
    cudaError_t e = cudaMemsetAsync(d_out, 0, (size_t)n * h * w, as_stream(stream));
    if (e != cudaSuccess) { ampis_set_error("polygon2mask memset: %s", cudaGetErrorString(e)); return AMPIS_ECUDA; }
    polygon2mask_kernel<<<n, 256, 0, as_stream(stream)>>>(d_xy, d_xy_off, h, w, d_out);
    AMPIS_CHECK_LAUNCH("polygon2mask_kernel");
    return AMPIS_OK;
}
