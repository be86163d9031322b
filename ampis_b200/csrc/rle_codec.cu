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
extern "C" int ampis_version(void) { return 201; }

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

// rleFrString, one THREAD per mask.  The format is a varint stream (5 payload bits per
// character, bit 0x20 = "more", bit 0x10 of the last character = sign) with a second-order delta: count m > 2 is
// stored relative to count m-2.  Only the low 32 bits of a number are kept, exactly what the reference's (uint) cast
// of its long keeps: characters beyond the 7th of a number contribute nothing.
//
// The format is serial by nature (a number's position depends on every character before it), and the serial walk is
// ~12 instructions per character for ONE thread: 32 masks per warp cost ~60 warp instructions per mask where the
// two warp-per-mask versions of this kernel (32 characters per step with a prefix sum, then 128 per step with
// ballots and shuffles) took 704 and ~450 -- the cooperative bookkeeping cost more than the work it shared.  What made
// the first thread-per-mask version of round 1 slow was its memory access, not its arithmetic; here
//   * characters are read as aligned 32-bit words (one load per four characters, funnel-shifted to the string's
//     alignment; a thread's loads walk one or two cache lines, so all but the first hit L1);
//   * counts leave in aligned 128-bit stores from a four-register window (scalar stores only for the few counts
//     before the first 16-byte boundary and after the last).
// Threads of a warp take consecutive masks: offsets, lengths and results are coalesced.
#define SD_THREADS 128

// state of one thread's walk; SD_CHAR consumes one character c = (char - 48) & 0x3f
struct SdState {
    u32 x, sh;            // the number being assembled and the shift of its next character
    u32 q1, q2, keep;     // counts m-1 and m-2 as the delta wants them (count 0 is nobody's base: keep = 0 until it is out)
    u32 b0, b1, b2, b3;   // the last four counts: one 128-bit store when they fill an aligned group
    int ph;               // counts emitted minus the counts before the first 16-byte boundary of the output
    u32 *op;              // where the next count goes
};

__device__ __forceinline__ void sd_emit(SdState &t, u32 v)
{
    if (t.ph < 0) {
        *t.op = v;                                   // below the first 16-byte boundary: the words before belong to the previous mask
    } else {
        t.b0 = t.b1; t.b1 = t.b2; t.b2 = t.b3; t.b3 = v;
        if ((t.ph & 3) == 3) *reinterpret_cast<uint4 *>(t.op - 3) = make_uint4(t.b0, t.b1, t.b2, t.b3);
    }
    t.ph++;
    t.op++;
}

// __funnelshift_lc(0, v, s) = v << s for s < 32 and 0 beyond: characters past the 7th of a number fall off, as in the
// reference's (uint) cast
__device__ __forceinline__ void sd_char(SdState &t, u32 c, bool force_end)
{
    t.x |= __funnelshift_lc(0u, c & 0x1fu, t.sh);
    t.sh += 5u;
    if (!(c & 0x20u) || force_end) {
        if (!(c & 0x20u) && (c & 0x10u)) t.x |= __funnelshift_lc(0u, 0xffffffffu, t.sh);
        const u32 v = t.x + t.q2;
        t.q2 = t.q1; t.q1 = v & t.keep; t.keep = 0xffffffffu;
        sd_emit(t, v);
        t.x = 0u; t.sh = 0u;
    }
}

__global__ void __launch_bounds__(SD_THREADS)
rle_string_decode_kernel(const uint8_t *__restrict__ chars, const i64 *__restrict__ chr_off, int n,
                         u32 *__restrict__ cnt, const i64 *__restrict__ cnt_off, int *__restrict__ cnt_len)
{
    const int i = (int)(blockIdx.x * (u32)SD_THREADS + threadIdx.x);
    if (i >= n) return;
    const i64 o0 = chr_off[i];
    const u32 len = (u32)min(chr_off[i + 1] - o0, (i64)0x7fffffff);
    const uint8_t *s = chars + o0;
    const u32 a = (u32)((uintptr_t)s & 3u);
    const u32 *sw = reinterpret_cast<const u32 *>(s - a);       // aligned words; only words holding a character are read
    const i64 co = cnt_off[i];
    SdState t;
    t.x = 0u; t.sh = 0u; t.q1 = 0u; t.q2 = 0u; t.keep = 0u;
    t.b0 = t.b1 = t.b2 = t.b3 = 0u;
    t.op = cnt + co;
    u32 *const out = t.op;
    // counts before the first 16-byte boundary of the output go out one by one; an unaligned count array gets scalar
    // stores throughout
    const int head = ((uintptr_t)cnt & 15u) == 0 ? (int)((0 - co) & 3) : 0x7fffffff;
    t.ph = -head;
    u32 wcur = len > 0u ? __ldg(sw) : 0u;
    const u32 groups = len >> 2;
    for (u32 q = 0; q < groups; q++) {                          // four characters at a time, all four present
        const u32 wnext = 4u * q + 4u - a < len ? __ldg(sw + q + 1) : 0u;
        const u32 bytes = __funnelshift_r(wcur, wnext, 8u * a);
        wcur = wnext;
        const u32 w6 = ((bytes & 0x7f7f7f7fu) + 0x10101010u) & 0x3f3f3f3fu;      // (char - 48) & 0x3f in every byte
        const bool last = 4u * q + 4u == len;                   // a string that stops in the middle of a number still yields it
        sd_char(t, w6 & 0xffu, false);
        sd_char(t, (w6 >> 8) & 0xffu, false);
        sd_char(t, (w6 >> 16) & 0xffu, false);
        sd_char(t, w6 >> 24, last);
    }
    const u32 tail = len & 3u;
    if (tail) {                                                 // the last one to three characters
        const u32 wnext = 4u * groups + 4u - a < len ? __ldg(sw + groups + 1) : 0u;
        const u32 bytes = __funnelshift_r(wcur, wnext, 8u * a);
        const u32 w6 = ((bytes & 0x7f7f7f7fu) + 0x10101010u) & 0x3f3f3f3fu;
        for (u32 b = 0; b < tail; b++) sd_char(t, (w6 >> (8u * b)) & 0xffu, b == tail - 1u);
    }
    const u32 m = (u32)(t.op - out);
    if (t.ph > 0) {                       // counts after the last full group
        const u32 left = (u32)t.ph & 3u;
        if (left >= 1u) t.op[-1] = t.b3;
        if (left >= 2u) t.op[-2] = t.b2;
        if (left == 3u) t.op[-3] = t.b1;
    }
    cnt_len[i] = (int)m;
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
    rle_string_decode_kernel<<<(unsigned)(((i64)n + SD_THREADS - 1) / SD_THREADS), SD_THREADS, 0, as_stream(stream)>>>(
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
