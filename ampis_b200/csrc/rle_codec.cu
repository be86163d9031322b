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
extern "C" int ampis_version(void) { return 100; }

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

// One thread per mask: the format is a byte-serial varint stream with a second-order
// delta (count i>2 is stored relative to count i-2), ~1.4 bytes per count and ~130
// counts per mask, so a mask is ~200 dependent byte steps.  A batch has 10^5..10^6
// masks, which is all the parallelism the chip needs; the strings are read through L1.
__global__ void __launch_bounds__(128)
rle_string_decode_kernel(const uint8_t *__restrict__ chars, const i64 *__restrict__ chr_off, int n,
                         u32 *__restrict__ cnt, const i64 *__restrict__ cnt_off, int *__restrict__ cnt_len)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *s = chars + chr_off[i];
    const i64 len = chr_off[i + 1] - chr_off[i];
    u32 *out = cnt + cnt_off[i];
    i64 p = 0;
    int m = 0;
    long long prev1 = 0, prev2 = 0;   // cnts[m-1], cnts[m-2]
    while (p < len) {
        long long x = 0;
        int k = 0;
        bool more = true;
        while (more && p < len) {
            long long c = (long long)s[p] - 48;
            if (k < 12) x |= (c & 0x1f) << (5 * k);
            more = (c & 0x20) != 0;
            p++;
            k++;
            if (!more && (c & 0x10) && k < 12) x |= (long long)(~0ULL << (5 * k));
        }
        if (m > 2) x += prev2;
        u32 v = (u32)x;
        out[m] = v;
        prev2 = prev1;
        prev1 = (long long)v;
        m++;
    }
    cnt_len[i] = m;
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
    rle_string_decode_kernel<<<(n + 127) / 128, 128, 0, as_stream(stream)>>>(d_chars, d_chr_off, n, d_cnt,
                                                                            d_cnt_off, d_cnt_len);
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
