// Dense intersection matrices by a TMA-staged, shared-memory tiled AND+popc kernel -- the third intersection kernel
// BASELINE.json's north_star names, built to be MEASURED next to the culled rows kernels and the tcgen05 contraction
// (profiles/crossover_r02.md), same quantity as both: I[r][c] = popcount(row AND col) for every pair of a group
// (what rleIou's run walk accumulates per pair, analyze.py:108,158; powder.py:82).
//
//   * operands: FULL-layout frames stored regularly (mask i at i * frame_chunks), described to the TMA unit by ONE
//     2-D tensor map {words of a frame, masks} (cuTensorMapEncodeTiled, 128-byte swizzle);
//   * a CTA owns a 64 x 64 tile of one group's matrix; a producer thread issues two cp.async.bulk.tensor.2d loads
//     per K step (64 row masks x 128 bytes, 64 column masks x 128 bytes) into a 4-stage ring, completion on
//     mbarriers (complete_tx), slots handed back by the consumer warps on a second set of mbarriers;
//   * 256 consumer threads hold 4 x 4 accumulators each and read the operands with 128-bit shared loads through the
//     swizzle (2-way conflicts at worst): 8 shared loads per 64 AND + 64 POPC;
//   * K range of a tile = the 128-byte steps where the union span of its rows meets the union span of its columns.
// The bound is the POPC pipe (16 lanes per clock and SM): G*P*H*W/32 word pairs per image whatever the masks look
// like -- which is why the culled kernels and the tensor-core contraction are the product paths.
#include <cuda.h>

#include "common.cuh"
#include "async.cuh"

#define TT_M 64                      // rows / columns of a tile
#define TT_KW 32                     // words per K step (128 bytes: one swizzle atom)
#define TT_STAGES 4
#define TT_CONSUMERS 256
#define TT_THREADS (TT_CONSUMERS + 32)
#define TT_STAGE_BYTES (2 * TT_M * TT_KW * 4)

__device__ __forceinline__ void tma_load_2d(u32 dst, const CUtensorMap *map, int c0, int c1, u32 bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__device__ __forceinline__ uint4 lds128(u32 addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

struct TmaTileArgs {
    const uint2 *span;
    const int *row_mask;
    const int *tile_grp, *tile_m0, *tile_n0;
    const int *grp_row_begin, *grp_row_count, *grp_col_begin, *grp_col_count;
    const i64 *grp_imat_off;
    int *imat;
    u32 frame_chunks;
};

__global__ void __launch_bounds__(TT_THREADS, 2)
intersect_tma_kernel(const __grid_constant__ CUtensorMap tmap, const TmaTileArgs p)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) u64 s_full[TT_STAGES], s_empty[TT_STAGES];
    __shared__ u32 s_lo[2], s_hi[2];
    // 1024-byte alignment of the stage ring (the swizzle pattern repeats every 1024 bytes)
    const u32 ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const u32 tid = threadIdx.x, lane = lane_id(), wid = tid >> 5;
    const int g = p.tile_grp[blockIdx.x], m0 = p.tile_m0[blockIdx.x], n0 = p.tile_n0[blockIdx.x];
    const int G = p.grp_row_count[g], P = p.grp_col_count[g];
    const int row0 = p.row_mask[p.grp_row_begin[g]] + m0;        // rows of a group are consecutive masks
    const int col0 = p.grp_col_begin[g] + n0;
    const int nr = min(TT_M, G - m0), nc = min(TT_M, P - n0);
    if (tid == 0) {
        for (int s = 0; s < TT_STAGES; s++) { mbar_init(smem_u32(&s_full[s]), 1); mbar_init(smem_u32(&s_empty[s]), TT_CONSUMERS / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_lo[0] = s_lo[1] = 0xffffffffu;
        s_hi[0] = s_hi[1] = 0u;
    }
    __syncthreads();
    // union span of the tile's rows and of its columns (128-bit chunks) -> K range in 128-byte steps
    if (tid < 2 * TT_M) {
        const int side = tid / TT_M, k = tid % TT_M;
        if (k < (side ? nc : nr)) {
            const uint2 sp = p.span[(side ? col0 : row0) + k];
            if (sp.y > sp.x) { atomicMin(&s_lo[side], sp.x); atomicMax(&s_hi[side], sp.y); }
        }
    }
    __syncthreads();
    const u32 lo = max(s_lo[0], s_lo[1]), hi = min(s_hi[0], s_hi[1]);
    const u32 k0 = lo / 8u, k1 = hi > lo ? (hi + 7u) / 8u : k0;                 // 8 chunks = 128 bytes per step
    const u32 nk = k1 > k0 ? k1 - k0 : 0u;

    u32 acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = 0u;

    if (wid == TT_CONSUMERS / 32) {
        // ---- producer: one thread feeds the ring
        if (lane == 0) {
            for (u32 k = 0; k < nk; k++) {
                const u32 s = k % TT_STAGES, round = k / TT_STAGES;
                if (round) mbar_wait(smem_u32(&s_empty[s]), (round - 1u) & 1u);
                const u32 full = smem_u32(&s_full[s]);
                mbar_arrive_tx(full, TT_STAGE_BYTES);
                const u32 dst = ring + s * TT_STAGE_BYTES;
                tma_load_2d(dst, &tmap, (int)((k0 + k) * TT_KW), row0, full);
                tma_load_2d(dst + TT_M * TT_KW * 4, &tmap, (int)((k0 + k) * TT_KW), col0, full);
            }
        }
    } else {
        // ---- consumers: thread (ty, tx) owns rows ty + 16 i and columns tx + 16 j
        const u32 tx = tid & 15u, ty = tid >> 4;
        for (u32 k = 0; k < nk; k++) {
            const u32 s = k % TT_STAGES;
            mbar_wait(smem_u32(&s_full[s]), (k / TT_STAGES) & 1u);
            const u32 a_base = ring + s * TT_STAGE_BYTES, b_base = a_base + TT_M * TT_KW * 4;
#pragma unroll 2
            for (u32 kc = 0; kc < 8; kc++) {
                uint4 a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const u32 r = ty + 16u * i, c = tx + 16u * i;
                    a[i] = lds128(a_base + r * 128u + ((kc ^ (r & 7u)) << 4));
                    b[i] = lds128(b_base + c * 128u + ((kc ^ (c & 7u)) << 4));
                }
#pragma unroll
                for (int i = 0; i < 4; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) acc[i][j] += popc_and(a[i], b[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&s_empty[s]));
        }
        // ---- epilogue: the tile's valid cells
        const i64 off = p.grp_imat_off[g];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int r = (int)ty + 16 * i;
            if (r >= nr) continue;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int c = (int)tx + 16 * j;
                if (c < nc) p.imat[off + (i64)(m0 + r) * P + n0 + c] = (int)acc[i][j];
            }
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

extern "C" int ampis_tma_tile(void) { return TT_M; }

extern "C" int ampis_intersect_tma(const void *d_bits, int64_t frame_chunks, int32_t n_masks, const uint32_t *d_span,
                                   const int32_t *d_row_mask, const int32_t *d_tile_grp, const int32_t *d_tile_m0,
                                   const int32_t *d_tile_n0, int32_t n_tiles, const int32_t *d_grp_row_begin,
                                   const int32_t *d_grp_row_count, const int32_t *d_grp_col_begin,
                                   const int32_t *d_grp_col_count, const int64_t *d_grp_imat_off, int32_t *d_imat,
                                   void *stream)
{
    AMPIS_REQUIRE(n_tiles >= 0 && n_masks >= 0 && frame_chunks > 0, "bad size");
    if (n_tiles == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_bits && d_span && d_row_mask && d_tile_grp && d_tile_m0 && d_tile_n0 && d_grp_row_begin &&
                      d_grp_row_count && d_grp_col_begin && d_grp_col_count && d_grp_imat_off && d_imat, "null pointer");
    AMPIS_REQUIRE(((uintptr_t)d_bits & 15u) == 0, "bits arena must be 16-byte aligned");
    AMPIS_REQUIRE(frame_chunks * 4 < ((int64_t)1 << 31), "frame too large for a 32-bit tensor coordinate");
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        const cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn) {
            ampis_set_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
            return AMPIS_ECUDA;
        }
        encode = (EncodeTiledFn)fn;
    }
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)frame_chunks * 4, (cuuint64_t)n_masks};
    const cuuint64_t strides[1] = {(cuuint64_t)frame_chunks * 16};
    const cuuint32_t box[2] = {TT_KW, TT_M};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(d_bits), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { ampis_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return AMPIS_ECUDA; }
    TmaTileArgs a;
    a.span = (const uint2 *)d_span; a.row_mask = d_row_mask; a.tile_grp = d_tile_grp; a.tile_m0 = d_tile_m0;
    a.tile_n0 = d_tile_n0; a.grp_row_begin = d_grp_row_begin; a.grp_row_count = d_grp_row_count;
    a.grp_col_begin = d_grp_col_begin; a.grp_col_count = d_grp_col_count; a.grp_imat_off = d_grp_imat_off;
    a.imat = d_imat; a.frame_chunks = (u32)frame_chunks;
    const size_t dyn = (size_t)TT_STAGES * TT_STAGE_BYTES + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        const cudaError_t e = cudaFuncSetAttribute(intersect_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) { ampis_set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return AMPIS_ECUDA; }
        attr_set = true;
    }
    intersect_tma_kernel<<<n_tiles, TT_THREADS, dyn, as_stream(stream)>>>(map, a);
    AMPIS_CHECK_LAUNCH("intersect_tma_kernel");
    return AMPIS_OK;
}
