// Packed masks -> RLE run counts (pycocotools rleEncode; reference call sites analyze.py:692,
// data_utils.py:275,423, structures.py:465, powder.py:209 via RLE.encode), and the pixel-class
// projection of analyze.seg_perf_iset (analyze.py:637-692).
//
// rleEncode walks the column-major pixels and emits alternating run lengths starting with a run of
// zeros (of length 0 when the first pixel is set).  On packed bits the run boundaries are the set
// bits of  t = x ^ ((x << 1) | last bit of the previous word)  with a virtual 0 before pixel 0:
// counts are the differences of consecutive boundary positions, closed by h*w.  Two passes per
// mask (count boundaries, then emit them in order through a block scan); FULL-layout masks only
// (bits beyond h*w in the last chunk are zero there).
#include "common.cuh"

#define ENC_THREADS 256

__device__ __forceinline__ u32 transitions(const u32 *__restrict__ w, i64 k, u64 nbits)
{
    // boundary bits of word k of a mask with nbits pixels (boundaries at positions >= nbits dropped)
    const u32 x = w[k];
    const u32 prev = k ? (w[k - 1] >> 31) : 0u;
    u32 t = x ^ ((x << 1) | prev);
    const u64 b0 = (u64)k * 32;
    if (b0 + 32 > nbits) t &= nbits > b0 ? ((1u << (u32)(nbits - b0)) - 1u) : 0u;
    return t;
}

__global__ void __launch_bounds__(ENC_THREADS)
rle_encode_count_kernel(const u32 *__restrict__ words, const i64 *__restrict__ bits_off,
                        const u32 *__restrict__ hh, const u32 *__restrict__ ww, int n, i64 *__restrict__ n_runs)
{
    __shared__ u32 s_sum;
    const int i = blockIdx.x;
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    const u64 nbits = (u64)hh[i] * ww[i];
    const i64 nwords = (i64)((nbits + 31) / 32);
    const u32 *w = words + bits_off[i] * 4;
    u32 acc = 0;
    for (i64 k = threadIdx.x; k < nwords; k += ENC_THREADS) acc += __popc(transitions(w, k, nbits));
    acc = warp_sum(acc);
    if (lane_id() == 0) atomicAdd(&s_sum, acc);
    __syncthreads();
    if (threadIdx.x == 0) n_runs[i] = (i64)s_sum + 1;
}

__global__ void __launch_bounds__(ENC_THREADS)
rle_encode_emit_kernel(const u32 *__restrict__ words, const i64 *__restrict__ bits_off,
                       const u32 *__restrict__ hh, const u32 *__restrict__ ww, int n,
                       const i64 *__restrict__ cnt_off, u32 *__restrict__ cnt, int *__restrict__ cnt_len)
{
    __shared__ u32 s_warp[ENC_THREADS / 32];
    __shared__ u32 s_base;
    const int i = blockIdx.x;
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const u64 nbits = (u64)hh[i] * ww[i];
    const i64 nwords = (i64)((nbits + 31) / 32);
    const u32 *w = words + bits_off[i] * 4;
    u32 *pos = cnt + cnt_off[i];               // boundary positions first, turned into counts below
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (i64 k0 = 0; k0 < nwords; k0 += ENC_THREADS) {
        const i64 k = k0 + threadIdx.x;
        u32 t = k < nwords ? transitions(w, k, nbits) : 0u;
        const u32 c = __popc(t);
        u32 incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const u32 v = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += v;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        u32 before = 0, tile_total = 0;
#pragma unroll
        for (u32 q = 0; q < ENC_THREADS / 32; q++) {
            const u32 v = s_warp[q];
            if (q < wid) before += v;
            tile_total += v;
        }
        u32 o = s_base + before + incl - c;
        while (t) {
            const u32 b = __ffs(t) - 1;
            t &= t - 1;
            pos[o++] = (u32)(k * 32) + b;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += tile_total;
        __syncthreads();
    }
    const u32 T = s_base;                      // number of boundaries; runs = T + 1
    __syncthreads();
    // positions -> counts, in place, back to front in tiles so no element is read after it is written
    // (element j needs j and j-1; process tiles from the end, reading both before any write of the tile)
    const u32 m = T + 1;
    for (i64 hi = (i64)m; hi > 0; hi -= ENC_THREADS) {
        const i64 j = hi - 1 - threadIdx.x;
        u32 v = 0;
        if (j >= 0) {
            const u32 end = j == (i64)T ? (u32)nbits : pos[j];
            const u32 start = j ? pos[j - 1] : 0u;
            v = end - start;
        }
        __syncthreads();
        if (j >= 0) pos[j] = v;
        __syncthreads();
    }
    if (threadIdx.x == 0) cnt_len[i] = (int)m;
}

extern "C" int ampis_bits_to_rle_count(const void *d_bits, const int64_t *d_bits_off, const uint32_t *d_h,
                                       const uint32_t *d_w, int32_t n, int64_t *d_n_runs, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits && d_bits_off && d_h && d_w && d_n_runs, "null pointer");
    rle_encode_count_kernel<<<n, ENC_THREADS, 0, as_stream(stream)>>>((const u32 *)d_bits, d_bits_off, d_h, d_w, n,
                                                                      d_n_runs);
    AMPIS_CHECK_LAUNCH("rle_encode_count_kernel");
    return AMPIS_OK;
}

extern "C" int ampis_bits_to_rle_emit(const void *d_bits, const int64_t *d_bits_off, const uint32_t *d_h,
                                      const uint32_t *d_w, int32_t n, const int64_t *d_cnt_off, uint32_t *d_cnt,
                                      int32_t *d_cnt_len, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits && d_bits_off && d_h && d_w && d_cnt_off && d_cnt && d_cnt_len, "null pointer");
    rle_encode_emit_kernel<<<n, ENC_THREADS, 0, as_stream(stream)>>>((const u32 *)d_bits, d_bits_off, d_h, d_w, n,
                                                                     d_cnt_off, d_cnt, d_cnt_len);
    AMPIS_CHECK_LAUNCH("rle_encode_emit_kernel");
    return AMPIS_OK;
}

// ---- pixel classes of matched pairs (analyze.seg_perf_iset) -------------------------------------
// tp |= g & p, fn |= g & ~p, fp |= ~g & p over all pairs, projected onto three frame bitmaps
// (np.logical_or.reduce over the matched pairs, analyze.py:645-652).  Warp per pair.
__global__ void __launch_bounds__(256)
project_pairs_kernel(const uint4 *__restrict__ bits, const i64 *__restrict__ bits_off,
                     const uint2 *__restrict__ reg, const uint2 *__restrict__ span,
                     const int *__restrict__ pair_gt, const int *__restrict__ pair_pr, int n_pairs,
                     u32 *__restrict__ tp, u32 *__restrict__ fn, u32 *__restrict__ fp)
{
    const int pair = (int)((blockIdx.x * (u32)blockDim.x + threadIdx.x) >> 5);
    if (pair >= n_pairs) return;
    const u32 lane = lane_id();
    const int g = pair_gt[pair], p = pair_pr[pair];
    const uint2 gs = span[g], ps = span[p];
    const uint4 *G = bits + bits_off[g] - reg[g].x, *P = bits + bits_off[p] - reg[p].x;
    const bool ge = gs.y > gs.x, pe = ps.y > ps.x;
    if (!ge && !pe) return;
    const u32 lo = ge && pe ? min(gs.x, ps.x) : (ge ? gs.x : ps.x);
    const u32 hi = ge && pe ? max(gs.y, ps.y) : (ge ? gs.y : ps.y);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (u32 c = lo + lane; c < hi; c += 32) {
        const uint4 a = (c >= gs.x && c < gs.y) ? ld_v4_nc(G + c) : z;
        const uint4 b = (c >= ps.x && c < ps.y) ? ld_v4_nc(P + c) : z;
        const u32 av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const u32 t = av[k] & bv[k], f = av[k] & ~bv[k], q = ~av[k] & bv[k];
            if (t) atomicOr(tp + 4 * (i64)c + k, t);
            if (f) atomicOr(fn + 4 * (i64)c + k, f);
            if (q) atomicOr(fp + 4 * (i64)c + k, q);
        }
    }
}

// class bitmaps from the three projections: mode 0 ('reduced') -> TP only, FN only, FP only, two or
// more; mode 1 ('all') -> codes 1..7 of TP + 2 FN + 4 FP (analyze.py:666-690).  out = n_out frames.
__global__ void __launch_bounds__(256)
class_masks_kernel(const u32 *__restrict__ tp, const u32 *__restrict__ fn, const u32 *__restrict__ fp,
                   i64 nwords, int mode, u32 *__restrict__ out)
{
    const i64 k = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nwords) return;
    const u32 T = tp[k], F = fn[k], P = fp[k];
    if (mode == 0) {
        out[k] = T & ~F & ~P;
        out[nwords + k] = ~T & F & ~P;
        out[2 * nwords + k] = ~T & ~F & P;
        out[3 * nwords + k] = (T & F) | (T & P) | (F & P);
    } else {
#pragma unroll
        for (u32 code = 1; code < 8; code++)
            out[(code - 1) * nwords + k] = ((code & 1u) ? T : ~T) & ((code & 2u) ? F : ~F) & ((code & 4u) ? P : ~P);
    }
}

extern "C" int ampis_project_pairs(const void *d_bits, const int64_t *d_bits_off, const uint32_t *d_reg,
                                   const uint32_t *d_span, const int32_t *d_pair_gt, const int32_t *d_pair_pr,
                                   int32_t n_pairs, int64_t frame_chunks, int32_t mode, void *d_tmp3,
                                   void *d_out_bits, void *stream)
{
    AMPIS_REQUIRE(n_pairs >= 0 && frame_chunks >= 0, "negative size");
    AMPIS_REQUIRE(mode == 0 || mode == 1, "bad mode");
    if (frame_chunks == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_tmp3 && d_out_bits, "null pointer");
    const i64 nwords = frame_chunks * 4;
    u32 *tp = (u32 *)d_tmp3, *fn = tp + nwords, *fp = fn + nwords;
    cudaError_t e = cudaMemsetAsync(d_tmp3, 0, (size_t)nwords * 12, as_stream(stream));
    if (e != cudaSuccess) { ampis_set_error("project memset: %s", cudaGetErrorString(e)); return AMPIS_ECUDA; }
    if (n_pairs) {
        AMPIS_REQUIRE(d_bits && d_bits_off && d_reg && d_span && d_pair_gt && d_pair_pr, "null pointer");
        project_pairs_kernel<<<(unsigned)(((i64)n_pairs * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
            (const uint4 *)d_bits, d_bits_off, (const uint2 *)d_reg, (const uint2 *)d_span, d_pair_gt, d_pair_pr,
            n_pairs, tp, fn, fp);
        AMPIS_CHECK_LAUNCH("project_pairs_kernel");
    }
    class_masks_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, as_stream(stream)>>>(tp, fn, fp, nwords, mode,
                                                                                       (u32 *)d_out_bits);
    AMPIS_CHECK_LAUNCH("class_masks_kernel");
    return AMPIS_OK;
}
