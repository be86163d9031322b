// Overlap of two AMPIS_LAYOUT_CROP windows and the AND+popc walk over it, shared by the crop rows kernels
// (intersect_crop.cu: every column scanned from shared memory; intersect_grid.cu: columns found through a
// uniform grid over the image).
#pragma once
#include "common.cuh"

struct Overlap {
    const u32 *A0, *B0;
    u32 nw, total, rn, cn;
};

__device__ __forceinline__ Overlap overlap_of(const u32 *A, const int4 rb, const u32 *B, const int4 cb)
{
    Overlap o;
    const u32 xa = (u32)max(rb.x, cb.x), xb = (u32)min(rb.z, cb.z);
    const u32 rw0 = (u32)rb.y >> 5, rw1 = (u32)rb.w >> 5, cw0 = (u32)cb.y >> 5, cw1 = (u32)cb.w >> 5;
    const u32 wa = max(rw0, cw0), wb = min(rw1, cw1);
    o.nw = wb - wa + 1u;
    o.total = (xb - xa + 1u) * o.nw;
    o.rn = rw1 - rw0 + 1u;
    o.cn = cw1 - cw0 + 1u;
    o.A0 = A + (xa - (u32)rb.x) * o.rn + (wa - rw0);
    o.B0 = B + (xa - (u32)cb.x) * o.cn + (wa - cw0);
    return o;
}

// popcount over the (column, band) pairs idx = first, first + stride, ... of an overlap
__device__ __forceinline__ u32 overlap_popc(const Overlap &o, u32 first, u32 stride)
{
    u32 dx = first / o.nw, dw = first - dx * o.nw;
    const u32 sdx = stride / o.nw, sdw = stride - sdx * o.nw;
    u32 acc = 0;
#pragma unroll 4
    for (u32 idx = first; idx < o.total; idx += stride) {
        acc += __popc(__ldg(o.A0 + dx * o.rn + dw) & __ldg(o.B0 + dx * o.cn + dw));
        dw += sdw; dx += sdx;
        if (dw >= o.nw) { dw -= o.nw; dx++; }
    }
    return acc;
}

