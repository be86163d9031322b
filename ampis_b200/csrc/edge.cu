// Boundary disagreement of matched mask pairs (analyze.mask_edge_distance, analyze.py:416-499).
//
// Reference: for every matched (gt, pred) pair decode both masks, crop to the merged box
// [r1:r2, c1:c2], list the false-positive pixels (pred & ~gt) and false-negative pixels (gt & ~pred)
// in np.where order (row-major) and take, for each, the smallest Euclidean distance to ANY pixel of
// the other mask inside the crop (an n x m distance table per pair).
//
// Here nothing is decoded: pixels are read straight from the packed column-major bit vectors, and
// the nearest pixel of a set to a point outside it is always a BOUNDARY pixel of the set (one with a
// 4-neighbour outside the set or outside the crop), so only boundary pixels are candidates.  Squared
// distances are exact integers; the result is sqrt in IEEE double, i.e. bit-identical to
// torch.sqrt(sum of squared differences) of the reference.
//
// Two launches, one CTA per pair: count (sizes of the four lists), then fill + distances.
#include "common.cuh"

#define EDGE_THREADS 256

struct EdgeArgs {
    const u32 *words;        // packed masks as 32-bit words
    const i64 *bits_off;
    const uint2 *reg;
    const uint2 *span;
    const u32 *h;
    const int *pair_gt, *pair_pr;
    const int4 *win;         // r1, r2, c1, c2 (half open, inside the frame)
};

struct MaskView {
    const u32 *W;            // biased: word k>>5 of the frame is W[k>>5]
    uint2 sp;
    u32 H;
    __device__ __forceinline__ bool at(int r, int c) const
    {
        const u64 k = (u64)c * H + (u32)r;
        const u32 ch = (u32)(k >> 7);
        if (ch < sp.x || ch >= sp.y) return false;
        return (W[k >> 5] >> (k & 31u)) & 1u;
    }
};

__device__ __forceinline__ MaskView view_of(const EdgeArgs &a, int m)
{
    MaskView v;
    v.W = a.words + (a.bits_off[m] - (i64)a.reg[m].x) * 4;
    v.sp = a.span[m];
    v.H = a.h[m];
    return v;
}

// pixel classes inside the window: bit0 FP, bit1 FN, bit2 gt boundary, bit3 pred boundary
__device__ __forceinline__ u32 classify(const MaskView &G, const MaskView &P, const int4 w, int r, int c)
{
    const bool g = G.at(r, c), p = P.at(r, c);
    u32 f = (p && !g ? 1u : 0u) | (g && !p ? 2u : 0u);
    if (g || p) {
        const bool up = r > w.x, dn = r + 1 < w.y, lf = c > w.z, rt = c + 1 < w.w;
        if (g) {
            const bool inner = up && dn && lf && rt && G.at(r - 1, c) && G.at(r + 1, c) && G.at(r, c - 1) && G.at(r, c + 1);
            if (!inner) f |= 4u;
        }
        if (p) {
            const bool inner = up && dn && lf && rt && P.at(r - 1, c) && P.at(r + 1, c) && P.at(r, c - 1) && P.at(r, c + 1);
            if (!inner) f |= 8u;
        }
    }
    return f;
}

__global__ void __launch_bounds__(EDGE_THREADS)
edge_count_kernel(const EdgeArgs a, int *__restrict__ counts)
{
    __shared__ u32 s_cnt[4];
    const int pair = blockIdx.x;
    if (threadIdx.x < 4) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int4 w = a.win[pair];
    const int Hw = max(w.y - w.x, 0), Ww = max(w.w - w.z, 0);
    const MaskView G = view_of(a, a.pair_gt[pair]), P = view_of(a, a.pair_pr[pair]);
    u32 c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    const i64 total = (i64)Hw * Ww;
    for (i64 idx = threadIdx.x; idx < total; idx += EDGE_THREADS) {
        const int r = w.x + (int)(idx / Ww), c = w.z + (int)(idx % Ww);
        const u32 f = classify(G, P, w, r, c);
        c0 += f & 1u; c1 += (f >> 1) & 1u; c2 += (f >> 2) & 1u; c3 += (f >> 3) & 1u;
    }
    c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); c3 = warp_sum(c3);
    if (lane_id() == 0) { atomicAdd(&s_cnt[0], c0); atomicAdd(&s_cnt[1], c1); atomicAdd(&s_cnt[2], c2); atomicAdd(&s_cnt[3], c3); }
    __syncthreads();
    if (threadIdx.x < 4) counts[pair * 4 + threadIdx.x] = (int)s_cnt[threadIdx.x];
}

// exclusive scan of one flag per thread over the CTA; returns this thread's offset, *total = CTA sum
__device__ __forceinline__ u32 block_scan_flag(bool flag, u32 *total, u32 *s_warp /*[EDGE_THREADS/32 + 1]*/)
{
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const u32 bal = __ballot_sync(0xffffffffu, flag);
    const u32 before = __popc(bal & ((1u << lane) - 1u));
    __syncthreads();                 // previous use of s_warp finished
    if (lane == 0) s_warp[wid] = __popc(bal);
    __syncthreads();
    u32 base = 0, tot = 0;
#pragma unroll
    for (u32 k = 0; k < EDGE_THREADS / 32; k++) {
        const u32 v = s_warp[k];
        if (k < wid) base += v;
        tot += v;
    }
    *total = tot;
    return base + before;
}

__global__ void __launch_bounds__(EDGE_THREADS)
edge_fill_kernel(const EdgeArgs a, const i64 *__restrict__ off_fp, const i64 *__restrict__ off_fn,
                 const i64 *__restrict__ off_gb, const i64 *__restrict__ off_pb, u32 *__restrict__ l_fp,
                 u32 *__restrict__ l_fn, u32 *__restrict__ l_gb, u32 *__restrict__ l_pb,
                 double *__restrict__ d_fp, double *__restrict__ d_fn, int *__restrict__ status)
{
    __shared__ u32 s_warp[EDGE_THREADS / 32 + 1];
    __shared__ u32 s_nb[2];
    __shared__ u32 s_tile[1024];
    const int pair = blockIdx.x;
    if (threadIdx.x < 2) s_nb[threadIdx.x] = 0;
    __syncthreads();
    const int4 w = a.win[pair];
    const int Hw = max(w.y - w.x, 0), Ww = max(w.w - w.z, 0);
    const MaskView G = view_of(a, a.pair_gt[pair]), P = view_of(a, a.pair_pr[pair]);
    u32 *fp = l_fp + off_fp[pair], *fn = l_fn + off_fn[pair], *gb = l_gb + off_gb[pair], *pb = l_pb + off_pb[pair];
    const i64 total = (i64)Hw * Ww;
    u32 nfp = 0, nfn = 0;
    for (i64 base = 0; base < total; base += EDGE_THREADS) {        // row-major order = np.where order
        const i64 idx = base + threadIdx.x;
        u32 f = 0, rc = 0;
        if (idx < total) {
            const int r = (int)(idx / Ww), c = (int)(idx % Ww);
            f = classify(G, P, w, w.x + r, w.z + c);
            rc = ((u32)r << 16) | (u32)c;
        }
        u32 t;
        const u32 pf = block_scan_flag(f & 1u, &t, s_warp);
        if (f & 1u) fp[nfp + pf] = rc;
        nfp += t;
        const u32 pn = block_scan_flag(f & 2u, &t, s_warp);
        if (f & 2u) fn[nfn + pn] = rc;
        nfn += t;
        if (f & 4u) gb[atomicAdd(&s_nb[0], 1u)] = rc;
        if (f & 8u) pb[atomicAdd(&s_nb[1], 1u)] = rc;
    }
    __syncthreads();
    const u32 ngb = s_nb[0], npb = s_nb[1];
    if (threadIdx.x == 0) status[pair] = ((nfp && !ngb) || (nfn && !npb)) ? 1 : 0;
    // distances: queries strided over threads, candidate boundary pixels staged through shared memory
    for (int which = 0; which < 2; which++) {
        const u32 nq = which ? nfn : nfp, nc = which ? npb : ngb;
        const u32 *q = which ? fn : fp, *cand = which ? pb : gb;
        double *out = (which ? d_fn + off_fn[pair] : d_fp + off_fp[pair]);
        if (!nq || !nc) continue;
        for (u32 q0 = 0; q0 < nq; q0 += EDGE_THREADS) {
            const u32 qi = q0 + threadIdx.x;
            const u32 me = qi < nq ? q[qi] : 0u;
            const int qr = (int)(me >> 16), qc = (int)(me & 0xffffu);
            u32 best = 0xffffffffu;
            for (u32 t0 = 0; t0 < nc; t0 += 1024) {
                const u32 tn = min(1024u, nc - t0);
                __syncthreads();
                for (u32 k = threadIdx.x; k < tn; k += EDGE_THREADS) s_tile[k] = cand[t0 + k];
                __syncthreads();
                if (qi < nq) {
                    for (u32 k = 0; k < tn; k++) {
                        const u32 v = s_tile[k];
                        const int dr = (int)(v >> 16) - qr, dc = (int)(v & 0xffffu) - qc;
                        best = min(best, (u32)(dr * dr + dc * dc));
                    }
                }
            }
            if (qi < nq) out[qi] = sqrt((double)best);
        }
    }
}

static int edge_args(EdgeArgs *a, const void *d_bits, const int64_t *d_bits_off, const uint32_t *d_reg,
                     const uint32_t *d_span, const uint32_t *d_h, const int32_t *d_pair_gt,
                     const int32_t *d_pair_pr, const int32_t *d_win)
{
    a->words = (const u32 *)d_bits; a->bits_off = d_bits_off; a->reg = (const uint2 *)d_reg;
    a->span = (const uint2 *)d_span; a->h = d_h; a->pair_gt = d_pair_gt; a->pair_pr = d_pair_pr;
    a->win = (const int4 *)d_win;
    return 0;
}

extern "C" int ampis_edge_count(const void *d_bits, const int64_t *d_bits_off, const uint32_t *d_reg,
                                const uint32_t *d_span, const uint32_t *d_h, const int32_t *d_pair_gt,
                                const int32_t *d_pair_pr, const int32_t *d_win, int32_t n_pairs,
                                int32_t *d_counts, void *stream)
{
    AMPIS_REQUIRE(n_pairs >= 0, "n_pairs < 0");
    if (n_pairs == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits && d_bits_off && d_reg && d_span && d_h && d_pair_gt && d_pair_pr && d_win && d_counts,
                  "null pointer");
    EdgeArgs a;
    edge_args(&a, d_bits, d_bits_off, d_reg, d_span, d_h, d_pair_gt, d_pair_pr, d_win);
    edge_count_kernel<<<n_pairs, EDGE_THREADS, 0, as_stream(stream)>>>(a, d_counts);
    AMPIS_CHECK_LAUNCH("edge_count_kernel");
    return AMPIS_OK;
}

extern "C" int ampis_edge_distances(const void *d_bits, const int64_t *d_bits_off, const uint32_t *d_reg,
                                    const uint32_t *d_span, const uint32_t *d_h, const int32_t *d_pair_gt,
                                    const int32_t *d_pair_pr, const int32_t *d_win, int32_t n_pairs,
                                    const int64_t *d_off_fp, const int64_t *d_off_fn, const int64_t *d_off_gb,
                                    const int64_t *d_off_pb, uint32_t *d_list_fp, uint32_t *d_list_fn,
                                    uint32_t *d_list_gb, uint32_t *d_list_pb, double *d_dist_fp,
                                    double *d_dist_fn, int32_t *d_status, void *stream)
{
    AMPIS_REQUIRE(n_pairs >= 0, "n_pairs < 0");
    if (n_pairs == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits && d_bits_off && d_reg && d_span && d_h && d_pair_gt && d_pair_pr && d_win && d_off_fp &&
                      d_off_fn && d_off_gb && d_off_pb && d_list_fp && d_list_fn && d_list_gb && d_list_pb &&
                      d_dist_fp && d_dist_fn && d_status, "null pointer");
    EdgeArgs a;
    edge_args(&a, d_bits, d_bits_off, d_reg, d_span, d_h, d_pair_gt, d_pair_pr, d_win);
    edge_fill_kernel<<<n_pairs, EDGE_THREADS, 0, as_stream(stream)>>>(
        a, d_off_fp, d_off_fn, d_off_gb, d_off_pb, d_list_fp, d_list_fn, d_list_gb, d_list_pb, d_dist_fp, d_dist_fn,
        d_status);
    AMPIS_CHECK_LAUNCH("edge_fill_kernel");
    return AMPIS_OK;
}
