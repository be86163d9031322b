// Matching summaries and histograms: the small reductions that follow the row kernel.
#include "common.cuh"

// One CTA per group (image).  AMPIS's matcher (analyze.py:166-174): a GT is a true positive
// when its best IoU is strictly above the threshold, otherwise a false negative; predictions
// never chosen by a true-positive GT are false positives (several GTs may share a prediction).
// The arg-max is threshold independent, so all thresholds are served by one row pass.
__global__ void __launch_bounds__(256)
match_counts_kernel(const int *__restrict__ best_col, const double *__restrict__ best_score,
                    const int *__restrict__ grp_row_begin, const int *__restrict__ grp_row_count,
                    const int *__restrict__ grp_col_count, const double *__restrict__ thresh, int n_thresh,
                    int *__restrict__ grp_counts, unsigned long long *__restrict__ totals)
{
    extern __shared__ u32 taken[];   // bitmap over the group's predictions
    __shared__ int s_tp, s_used;
    const int g = blockIdx.x;
    const int r0 = grp_row_begin[g], G = grp_row_count[g], P = grp_col_count[g];
    const int words = (P + 31) / 32;
    for (int t = 0; t < n_thresh; t++) {
        const double th = thresh[t];
        for (int k = threadIdx.x; k < words; k += blockDim.x) taken[k] = 0u;
        if (threadIdx.x == 0) { s_tp = 0; s_used = 0; }
        __syncthreads();
        int tp = 0;
        for (int k = threadIdx.x; k < G; k += blockDim.x) {
            const int c = best_col[r0 + k];
            if (best_score[r0 + k] > th && c >= 0) {
                tp++;
                atomicOr(&taken[c >> 5], 1u << (c & 31));
            }
        }
        tp = (int)warp_sum((u32)tp);
        if (lane_id() == 0 && tp) atomicAdd(&s_tp, tp);
        __syncthreads();
        int used = 0;
        for (int k = threadIdx.x; k < words; k += blockDim.x) used += __popc(taken[k]);
        used = (int)warp_sum((u32)used);
        if (lane_id() == 0 && used) atomicAdd(&s_used, used);
        __syncthreads();
        if (threadIdx.x == 0) {
            const int TP = s_tp, FP = P - s_used, FN = G - s_tp;
            int *o = grp_counts + ((i64)g * n_thresh + t) * 3;
            o[0] = TP; o[1] = FP; o[2] = FN;
            if (totals) {
                atomicAdd(totals + 3 * t + 0, (unsigned long long)TP);
                atomicAdd(totals + 3 * t + 1, (unsigned long long)FP);
                atomicAdd(totals + 3 * t + 2, (unsigned long long)FN);
            }
        }
        __syncthreads();
    }
}

extern "C" int ampis_match_counts(const int32_t *d_best_col, const double *d_best_score,
                                  const int32_t *d_grp_row_begin, const int32_t *d_grp_row_count,
                                  const int32_t *d_grp_col_count, int32_t n_groups, int32_t max_cols,
                                  const double *d_thresh, int32_t n_thresh, int32_t *d_grp_counts,
                                  int64_t *d_totals, void *stream)
{
    AMPIS_REQUIRE(n_groups >= 0 && n_thresh >= 0 && max_cols >= 0, "negative size");
    if (n_groups == 0 || n_thresh == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_best_col && d_best_score && d_grp_row_begin && d_grp_row_count && d_grp_col_count &&
                      d_thresh && d_grp_counts, "null pointer");
    const size_t smem = (size_t)((max_cols + 31) / 32 + 1) * sizeof(u32);
    AMPIS_REQUIRE(smem <= 200 * 1024, "too many columns per group for the shared bitmap");
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(match_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    match_counts_kernel<<<n_groups, 256, smem, as_stream(stream)>>>(
        d_best_col, d_best_score, d_grp_row_begin, d_grp_row_count, d_grp_col_count, d_thresh, n_thresh,
        d_grp_counts, (unsigned long long *)d_totals);
    AMPIS_CHECK_LAUNCH("match_counts_kernel");
    return AMPIS_OK;
}

// Satellite assignment (powder.py:85-96): satellite s goes to particle argmax when
// I/area(s) > thresh (NaN when area(s)==0 => unmatched).  Per particle the number of owned
// satellites feeds the satellites-per-particle histogram (powder.py:529-542).
__global__ void __launch_bounds__(256)
satellite_counts_kernel(const int *__restrict__ best_col, const u32 *__restrict__ best_inter,
                        const u32 *__restrict__ area, const int *__restrict__ row_mask,
                        const int *__restrict__ grp_row_begin, const int *__restrict__ grp_row_count,
                        const int *__restrict__ grp_col_count, double thresh, int *__restrict__ grp_counts,
                        unsigned long long *__restrict__ spp_hist, int n_bins)
{
    extern __shared__ u32 owned[];   // satellites per particle of this group
    __shared__ int s_matched, s_particles;
    const int g = blockIdx.x;
    const int r0 = grp_row_begin[g], S = grp_row_count[g], P = grp_col_count[g];
    for (int k = threadIdx.x; k < P; k += blockDim.x) owned[k] = 0u;
    if (threadIdx.x == 0) { s_matched = 0; s_particles = 0; }
    __syncthreads();
    int matched = 0;
    for (int k = threadIdx.x; k < S; k += blockDim.x) {
        const u32 a = area[row_mask[r0 + k]];
        const int c = best_col[r0 + k];
        const double score = (double)best_inter[r0 + k] / (double)a;
        if (score > thresh && c >= 0) {
            matched++;
            atomicAdd(&owned[c], 1u);
        }
    }
    matched = (int)warp_sum((u32)matched);
    if (lane_id() == 0 && matched) atomicAdd(&s_matched, matched);
    __syncthreads();
    int np = 0;
    for (int k = threadIdx.x; k < P; k += blockDim.x) {
        const u32 o = owned[k];
        if (o) {
            np++;
            if (spp_hist && n_bins > 0) atomicAdd(spp_hist + min((int)o, n_bins - 1), 1ull);
        }
    }
    np = (int)warp_sum((u32)np);
    if (lane_id() == 0 && np) atomicAdd(&s_particles, np);
    __syncthreads();
    if (threadIdx.x == 0) {
        int *o = grp_counts + (i64)g * 4;
        o[0] = s_matched; o[1] = S - s_matched; o[2] = s_particles; o[3] = P;
    }
}

extern "C" int ampis_satellite_counts(const int32_t *d_best_col, const uint32_t *d_best_inter,
                                      const uint32_t *d_area, const int32_t *d_row_mask,
                                      const int32_t *d_grp_row_begin, const int32_t *d_grp_row_count,
                                      const int32_t *d_grp_col_count, int32_t n_groups, int32_t max_cols,
                                      double thresh, int32_t *d_grp_counts, int64_t *d_spp_hist, int32_t n_bins,
                                      void *stream)
{
    AMPIS_REQUIRE(n_groups >= 0 && max_cols >= 0, "negative size");
    if (n_groups == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_best_col && d_best_inter && d_area && d_row_mask && d_grp_row_begin && d_grp_row_count &&
                      d_grp_col_count && d_grp_counts, "null pointer");
    const size_t smem = (size_t)(max_cols + 1) * sizeof(u32);
    AMPIS_REQUIRE(smem <= 200 * 1024, "too many particles per image for the shared counters");
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(satellite_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    satellite_counts_kernel<<<n_groups, 256, smem, as_stream(stream)>>>(
        d_best_col, d_best_inter, d_area, d_row_mask, d_grp_row_begin, d_grp_row_count, d_grp_col_count, thresh,
        d_grp_counts, (unsigned long long *)d_spp_hist, n_bins);
    AMPIS_CHECK_LAUNCH("satellite_counts_kernel");
    return AMPIS_OK;
}

// Uniform-bin histogram with per-CTA shared-memory privatisation.
__global__ void __launch_bounds__(256)
hist_u32_kernel(const u32 *__restrict__ v, i64 n, u32 lo, u32 bw, unsigned long long *__restrict__ hist,
                int n_bins)
{
    extern __shared__ u32 h[];
    for (int k = threadIdx.x; k < n_bins; k += blockDim.x) h[k] = 0u;
    __syncthreads();
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const u32 x = v[i];
        int b = x <= lo ? 0 : (int)min((u64)(x - lo) / bw, (u64)(n_bins - 1));
        atomicAdd(&h[b], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n_bins; k += blockDim.x)
        if (h[k]) atomicAdd(hist + k, (unsigned long long)h[k]);
}

extern "C" int ampis_hist_u32(const uint32_t *d_values, int64_t n, uint32_t lo, uint32_t bin_width,
                              int64_t *d_hist, int32_t n_bins, void *stream)
{
    AMPIS_REQUIRE(n >= 0 && n_bins > 0 && bin_width > 0, "bad size");
    AMPIS_REQUIRE(n_bins <= 12288, "at most 12288 bins");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_values && d_hist, "null pointer");
    int sms = ampis_sm_count();
    if (sms <= 0) return AMPIS_ECUDA;
    i64 blocks = (n + 256 * 16 - 1) / (256 * 16);
    if (blocks > (i64)sms * 8) blocks = (i64)sms * 8;
    hist_u32_kernel<<<(unsigned)blocks, 256, (size_t)n_bins * sizeof(u32), as_stream(stream)>>>(
        d_values, n, lo, bin_width, (unsigned long long *)d_hist, n_bins);
    AMPIS_CHECK_LAUNCH("hist_u32_kernel");
    return AMPIS_OK;
}
