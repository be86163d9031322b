// COCO compressed-RLE string codec on the GPU, plus library-wide error state.
// Replaces pycocotools rleFrString / rleToString, which the
// reference re-runs on every RLE.iou / merge / area call (analyze.py:158, powder.py:82).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void ampis_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char *ampis_last_error(void) { return g_err; }
extern "C" int ampis_version(void) { return 200; }

extern "C" int ampis_sm_count(void)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
        ampis_set_error("ampis_sm_count: %s", cudaGetErrorString(cudaGetLastError()));
        return AMPIS_ECUDA;
    }
    return sms;
}

// rleFrString, one WARP per mask, 32 characters per step.  The format is a
// varint stream (5 payload bits per character, bit 0x20 = "more", bit 0x10 of the last character =
// sign) with a second-order delta: count m > 2 is stored relative to count m-2.
//   1. a ballot of the "last character of a number" flags splits the 32 characters into numbers;
//   2. every character contributes payload << 5*(its position inside its number) (nothing beyond the 7th
//      character: only the low 32 bits are kept, exactly what the reference's (uint) cast of its long
//      keeps); the contributions of a number fill disjoint bit ranges, so its value is a SUM over its
//      lanes = a difference of two warp prefix sums (exact modulo 2^32);
//   3. numbers are compacted to the low lanes through shared memory (slot = rank of the end flag) and the
//      delta is undone by a stride-2 warp scan (two interleaved chains; count 0 belongs to no chain:
//      count 2 is stored absolute);
//   4. a number cut by the 32-character boundary carries its partial value into the next step.
// Reads and writes are coalesced; the previous thread-per-mask version spent its time in
// 32-way divergent byte loads.
__global__ void __launch_bounds__(256)
rle_string_decode_kernel(const uint8_t *__restrict__ chars, const i64 *__restrict__ chr_off, int n,
                         u32 *__restrict__ cnt, const i64 *__restrict__ cnt_off, int *__restrict__ cnt_len)
{
    __shared__ u32 s_x[8][32];            // numbers of a step, compacted (one row per warp)
    const int i = (int)((blockIdx.x * (u32)blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const u32 lane = lane_id(), wid = threadIdx.x >> 5;
    const u32 lt = (1u << lane) - 1u;
    const uint8_t *s = chars + chr_off[i];
    const i64 len = chr_off[i + 1] - chr_off[i];
    u32 *out = cnt + cnt_off[i];
    u32 m_base = 0;                       // numbers emitted so far
    u32 carry_x = 0, carry_k = 0;         // partial number continued from the previous step
    u32 carry_e = 0, carry_o = 0;         // last value of the even / odd delta chain (count 0 excluded)
    for (i64 p0 = 0; p0 < len; p0 += 32) {
        const i64 pos = p0 + lane;
        const bool valid = pos < len;
        const u32 c = valid ? (u32)s[pos] - 48u : 0u;
        // a string that stops in the middle of a number still yields that number (as the serial code does)
        const bool end = valid && (!(c & 0x20u) || pos == len - 1);
        const u32 e_mask = __ballot_sync(0xffffffffu, end);
        const u32 below = e_mask & lt;
        const u32 first = below ? 32u - (u32)__clz(below) : 0u;             // first lane of my number in this step
        const u32 k = lane - first + (below ? 0u : carry_k);                // my position inside my number
        const u32 contrib = (valid && k < 7u) ? (c & 0x1fu) << (5u * k) : 0u;
        // the characters of a number fill disjoint bit ranges, so OR-ing them is adding them, and a sum over lanes
        // first..me is a difference of two prefix sums (exact modulo 2^32, which is all the reference keeps)
        u32 P = contrib;
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, P, d);
            if (lane >= d) P += t;
        }
        const u32 before = __shfl_sync(0xffffffffu, P, (first + 31u) & 31u);      // P[first - 1]
        u32 x = P - (first ? before : 0u);
        if (!below) x |= carry_x;
        if (end && !(c & 0x20u) && (c & 0x10u)) {
            const u32 sh = 5u * (k + 1u);
            if (sh < 32u) x |= 0xffffffffu << sh;
        }
        const u32 nn = __popc(e_mask);
        // number t of this step: written to slot t by the lane that ends it, read back by lane t
        if (end) s_x[wid][__popc(below)] = x;
        __syncwarp();
        const u32 m = m_base + lane;
        const bool have = lane < nn;
        const u32 xv = have ? s_x[wid][lane] : 0u;
        __syncwarp();
        u32 y = (have && m != 0) ? xv : 0u;
#pragma unroll
        for (u32 d = 2; d < 32; d <<= 1) {
            const u32 t = __shfl_up_sync(0xffffffffu, y, d);
            if (lane >= d) y += t;
        }
        const u32 w = y + ((m & 1u) ? carry_o : carry_e);
        if (have) out[m] = m == 0 ? xv : w;
        // chains continue from the last two numbers of this step
        if (nn >= 1) {
            const u32 w1 = __shfl_sync(0xffffffffu, w, nn - 1);
            const u32 w2 = __shfl_sync(0xffffffffu, w, nn >= 2 ? nn - 2 : 0);
            const u32 m1 = m_base + nn - 1;
            if (m1 & 1u) { carry_o = w1; if (nn >= 2) carry_e = w2; }
            else { carry_e = w1; if (nn >= 2) carry_o = w2; }
        }
        m_base += nn;
        // characters above the last end flag start a number that finishes in a later step
        const int last_end = e_mask ? 31 - __clz(e_mask) : -1;
        const u32 total = __shfl_sync(0xffffffffu, P, 31);
        const u32 upto = last_end >= 0 ? __shfl_sync(0xffffffffu, P, last_end) : 0u;
        carry_x = (e_mask ? 0u : carry_x) | (total - upto);
        carry_k = (e_mask ? 0u : carry_k) + (u32)(31 - last_end);
    }
    if (lane == 0) cnt_len[i] = (int)m_base;
}

__global__ void __launch_bounds__(128)
rle_string_encode_kernel(const u32 *__restrict__ cnt, const i64 *__restrict__ cnt_off,
                         const int *__restrict__ cnt_len, int n, uint8_t *__restrict__ chars,
                         const i64 *__restrict__ chr_off, int *__restrict__ chr_len)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const u32 *c = cnt + cnt_off[i];
    const int m = cnt_len[i];
    uint8_t *s = chars + chr_off[i];
    int p = 0;
    long long prev1 = 0, prev2 = 0;
    for (int j = 0; j < m; j++) {
        long long v = (long long)c[j];
        long long x = v;
        if (j > 2) x -= prev2;
        prev2 = prev1;
        prev1 = v;
        bool more = true;
        while (more) {
            int ch = (int)(x & 0x1f);
            x >>= 5;
            more = (ch & 0x10) ? x != -1 : x != 0;
            if (more) ch |= 0x20;
            s[p++] = (uint8_t)(ch + 48);
        }
    }
    chr_len[i] = p;
}

extern "C" int ampis_rle_string_decode(const uint8_t *d_chars, const int64_t *d_chr_off, int32_t n,
                                       uint32_t *d_cnt, const int64_t *d_cnt_off, int32_t *d_cnt_len,
                                       void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_chars && d_chr_off && d_cnt && d_cnt_off && d_cnt_len, "null pointer");
    rle_string_decode_kernel<<<(unsigned)(((i64)n * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(
        d_chars, d_chr_off, n, d_cnt, d_cnt_off, d_cnt_len);
    AMPIS_CHECK_LAUNCH("rle_string_decode_kernel");
    return AMPIS_OK;
}

extern "C" int ampis_rle_string_encode(const uint32_t *d_cnt, const int64_t *d_cnt_off,
                                       const int32_t *d_cnt_len, int32_t n, uint8_t *d_chars,
                                       const int64_t *d_chr_off, int32_t *d_chr_len, void *stream)
{
    AMPIS_REQUIRE(n >= 0, "n < 0");
    if (n == 0) return AMPIS_OK;
    AMPIS_REQUIRE(d_chars && d_chr_off && d_cnt && d_cnt_off && d_cnt_len && d_chr_len, "null pointer");
    rle_string_encode_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(d_cnt, d_cnt_off, d_cnt_len, n,
                                                                            d_chars, d_chr_off, d_chr_len);
    AMPIS_CHECK_LAUNCH("rle_string_encode_kernel");
    return AMPIS_OK;
}
