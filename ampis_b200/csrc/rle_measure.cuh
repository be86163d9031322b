// Warp-level measurement of one RLE mask, shared by rle_measure_kernel and the fused
// measure+paint kernel.
#pragma once
#include "common.cuh"

struct MaskMeasure {
    u32 area;     // number of 1-pixels (rleArea)
    u32 first;    // first 1-pixel (column-major index), 0xffffffff if none
    u32 last;     // one past the last 1-pixel
    u32 ymin, ymax;
    u64 total;    // sum of all run counts
};

// One group of L lanes (a warp by default) walks the m run counts of a mask 4*L at a time: every lane takes FOUR consecutive runs (two
// 0-runs and two 1-runs, so all lanes do the same work), sums them locally, and ONE warp scan of the lane
// totals places them -- a typical mask (~75 runs) costs one scan instead of three and keeps 19 lanes busy
// on the per-run statistics instead of 16 lanes three times.
// Run end positions are written to cum_s[j] for j < cum_s_cap (shared memory, may be null)
// and to cum_g[j] (global memory, may be null).  Result valid in all lanes of the group.  Small masks
// (30-75 runs) leave most of a warp idle: L = 8 or 16 puts four or two masks on a warp.
template <int L = 32>
__device__ __forceinline__ MaskMeasure warp_measure(const u32 *__restrict__ cnt, int m, u32 H, u64 HW,
                                                    u32 *cum_s, int cum_s_cap, u32 *cum_g)
{
    const u32 lane = lane_id() & (u32)(L - 1);
    const u32 gm = group_mask<L>();
    const FastDiv byH = fastdiv_make(H);
    u64 carry = 0;
    u32 a = 0, first = 0xffffffffu, last = 0, ymin = 0xffffffffu, ymax = 0;
    for (int jb = 0; jb < m; jb += 4 * L) {
        const int j0 = jb + 4 * (int)lane;
        u32 c[4];
#pragma unroll
        for (int q = 0; q < 4; q++) c[q] = j0 + q < m ? __ldg(cnt + j0 + q) : 0u;
        u64 loc[4];
        loc[0] = c[0];
#pragma unroll
        for (int q = 1; q < 4; q++) loc[q] = loc[q - 1] + c[q];
        u64 incl = loc[3];
#pragma unroll
        for (int d = 1; d < L; d <<= 1) {
            const u64 t = __shfl_up_sync(gm, incl, d, L);
            if ((int)lane >= d) incl += t;
        }
        const u64 base = carry + incl - loc[3];
        u32 e[4];
#pragma unroll
        for (int q = 0; q < 4; q++) e[q] = (u32)min(base + loc[q], (u64)0xffffffffu);
        // run ends of the lane: one 128-bit shared store when the four are wanted (scalar stores of a lane
        // stride of four words would be 4-way bank conflicts)
        const bool vec = cum_s && j0 + 3 < min(m, cum_s_cap) && ((uintptr_t)(cum_s + j0) & 15u) == 0;
        if (vec) *reinterpret_cast<uint4 *>(cum_s + j0) = make_uint4(e[0], e[1], e[2], e[3]);
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = j0 + q;
            const u64 end64 = base + loc[q];
            const u32 end = e[q];
            if (j < m) {
                if (cum_s && !vec && j < cum_s_cap) cum_s[j] = end;
                if (cum_g) cum_g[j] = end;
            }
            if ((q & 1) && c[q] > 0 && j < m && end64 <= HW) {         // jb and 4*lane are even: odd q = 1-run
                const u32 start = end - c[q];
                a += c[q];
                first = min(first, start);
                last = max(last, end);
                const u32 xs = fastdiv(start, byH), xe = fastdiv(end - 1, byH);
                if (xs != xe) { ymin = 0; ymax = H - 1; }
                else { ymin = min(ymin, start - xs * H); ymax = max(ymax, end - 1 - xe * H); }
            }
        }
        carry = __shfl_sync(gm, carry + incl, L - 1, L);
    }
    MaskMeasure r;
    r.area = group_sum<L>(a, gm);
    r.first = group_min<L>(first, gm);
    r.last = group_max<L>(last, gm);
    r.ymin = group_min<L>(ymin, gm);
    r.ymax = group_max<L>(ymax, gm);
    r.total = carry;
    return r;
}

// AMPIS_LAYOUT_CROP window of a mask with tight box bb: columns bb.x..bb.z, and for each column
// the absolute 32-row bands (bb.y >> 5)..(bb.w >> 5).  Number of 32-bit words stored.
__device__ __forceinline__ u32 crop_words(const int4 bb)
{
    if (bb.z < bb.x) return 0u;
    return (u32)(bb.z - bb.x + 1) * (u32)((bb.w >> 5) - (bb.y >> 5) + 1);
}

// Derived per-mask records written by lane 0.
__device__ __forceinline__ void store_measure(const MaskMeasure &ms, u32 H, u64 HW, int layout, int i,
                                              u32 *area, int *bbox, u32 *span, u32 *reg, int *status,
                                              uint2 *span_out, uint2 *reg_out, int4 *bbox_out = nullptr)
{
    const u32 nchunks = (u32)((HW + AMPIS_CHUNK_BITS - 1) / AMPIS_CHUNK_BITS);
    u32 slo = 0, shi = 0;
    int4 bb = make_int4(0, 0, -1, -1);
    if (ms.area > 0) {
        slo = ms.first / AMPIS_CHUNK_BITS;
        shi = min((ms.last + AMPIS_CHUNK_BITS - 1) / AMPIS_CHUNK_BITS, nchunks);
        bb = make_int4((int)(ms.first / H), (int)ms.ymin, (int)((ms.last - 1) / H), (int)ms.ymax);
    }
    const uint2 sp = make_uint2(slo, shi);
    uint2 rg = layout == AMPIS_LAYOUT_FULL ? make_uint2(0u, nchunks) : sp;
    if (layout == AMPIS_LAYOUT_CROP) rg = make_uint2(0u, (crop_words(bb) + 3u) / 4u);
    area[i] = ms.area;
    reinterpret_cast<int4 *>(bbox)[i] = bb;
    reinterpret_cast<uint2 *>(span)[i] = sp;
    reinterpret_cast<uint2 *>(reg)[i] = rg;
    status[i] = ms.total == HW ? 0 : AMPIS_ST_BAD_TOTAL;
    *span_out = sp;
    *reg_out = rg;
    if (bbox_out) *bbox_out = bb;
}
