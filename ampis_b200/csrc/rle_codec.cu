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

// ---- the same decoder for SMALL batches (one image: a few hundred strings), where the time is one thread's latency
// and not throughput: a WARP per mask, four characters per lane, 128 per step (the second of the four formulations of
// profiles/experiments_r02.md; 160 us per 250,000 masks against 128 for the thread-per-mask walk, but one string is
// decoded in ~2 us instead of ~12).
//   1. a lane loads its four characters as two aligned words + a funnel shift and walks them serially: numbers that
//      end inside the lane are complete except the first one, which may have begun in an earlier lane (or step);
//   2. that beginning is at most the tail of the nearest lower lane holding a number end and the two lanes after it
//      (3 + 4 + 4 characters cover the 7 that matter): one ballot and three shuffles, no scan;
//   3. numbers are compacted through shared memory (slot = numbers ending in lower lanes, from three ballots of the
//      bits of the per-lane count) and read back four consecutive numbers per lane;
//   4. the delta is undone by two warp scans (even / odd chain sums of a lane's four numbers; count 0 belongs to no
//      chain: count 2 is stored absolute);
//   5. an unfinished number and the two chain values carry into the next step (strings of more than 128 characters).
#define SD_SMALL_BATCH 16384        // masks up to which the warp-per-mask form is launched
#define SD_WARPS 8

__device__ __forceinline__ u32 shl32(u32 x, u32 s) { return s < 32u ? x << s : 0u; }

// The characters between the last number end below virtual lane L (0..32) and lane L: value of the first seven
// (pv) and how many they are (pl, saturated at 8).  pk = a lane's characters after its last number end (all four if
// it holds no end): value in bits 0..19, count in bits 24..26.  E = lanes holding a number end.
__device__ __forceinline__ void sd_incoming(u32 E, u32 L, u32 pk, u32 carry_x, u32 carry_k, u32 &pv, u32 &pl)
{
    const u32 below = L >= 32u ? E : (E & ((1u << L) - 1u));
    const int pe = below ? 31 - __clz(below) : -1;
    const u32 g0 = __shfl_sync(0xffffffffu, pk, max(pe, 0));
    const u32 g1 = __shfl_sync(0xffffffffu, pk, min(pe + 1, 31));
    const u32 g2 = __shfl_sync(0xffffffffu, pk, min(pe + 2, 31));
    if (pe >= 0) { pv = g0 & 0xffffffu; pl = g0 >> 24; }
    else { pv = carry_x; pl = carry_k; }
    if (pe + 1 < (int)L) { pv |= shl32(g1 & 0xffffffu, 5u * pl); pl += g1 >> 24; }
    if (pe + 2 < (int)L) { pv |= shl32(g2 & 0xffffffu, 5u * pl); pl += g2 >> 24; }
    if (pe + 3 < (int)L) pl = 8u;                    // 3 + 4 + 4 or more characters already
    pl = min(pl, 8u);
}

__global__ void __launch_bounds__(SD_WARPS * 32)
rle_string_decode_warp_kernel(const uint8_t *__restrict__ chars, const i64 *__restrict__ chr_off, int n,
                         u32 *__restrict__ cnt, const i64 *__restrict__ cnt_off, int *__restrict__ cnt_len)
{
    __shared__ __align__(16) u32 s_x[SD_WARPS][128];            // numbers of a step, compacted (one row per warp)
    const int i = (int)((blockIdx.x * (u32)blockDim.x + threadIdx.x) >> 5);
    if (i >= n) return;
    const u32 lane = lane_id();
    u32 *S = s_x[threadIdx.x >> 5];
    const u32 lt = (1u << lane) - 1u;
    const uint8_t *s = chars + chr_off[i];
    const i64 len = chr_off[i + 1] - chr_off[i];
    const u32 a = (u32)((uintptr_t)s & 3u);
    const u32 *sw = reinterpret_cast<const u32 *>(s - a);       // aligned words; only words holding a character are read
    u32 *out = cnt + cnt_off[i];
    u32 m_base = 0;                       // numbers emitted so far
    u32 carry_x = 0, carry_k = 0;         // unfinished number continued from the previous step (count saturated at 8)
    u32 carry_e = 0, carry_o = 0;         // last value of the even / odd delta chain (count 0 excluded)
    for (i64 p0 = 0; p0 < len; p0 += 128) {
        const i64 rem = len - (p0 + 4 * (i64)lane);             // characters from my first one to the end of the string
        const u32 nv = rem <= 0 ? 0u : (rem >= 4 ? 4u : (u32)rem);
        const u32 lastb = (rem >= 1 && rem <= 4) ? nv - 1u : 8u;   // the string's last character, if it is one of mine
        u32 bytes = 0;
        if (nv) {
            const u32 *wp = sw + ((p0 >> 2) + lane);
            const u32 w0 = __ldg(wp);
            const u32 w1 = a + nv > 4u ? __ldg(wp + 1) : 0u;
            bytes = __funnelshift_r(w0, w1, 8u * a);
        }
        // ---- my four characters: numbers that end here (the first one may have begun earlier: finished below)
        u32 cur = 0, curk = 0, nloc = 0, hk = 0;
        u32 v0 = 0, v1 = 0, v2 = 0, v3 = 0;
        bool hs = false;
#pragma unroll
        for (u32 b = 0; b < 4; b++) {
            const u32 c = ((bytes >> (8u * b)) & 0xffu) - 48u;
            if (b < nv) {
                cur |= (c & 0x1fu) << (5u * curk);
                curk++;
                // a string that stops in the middle of a number still yields that number
                if (!(c & 0x20u) || b == lastb) {
                    const bool sign = !(c & 0x20u) && (c & 0x10u);
                    u32 x = cur;
                    if (nloc == 0) { hk = curk; hs = sign; }
                    else if (sign) x |= 0xffffffffu << (5u * curk);
                    if (nloc == 0) v0 = x; else if (nloc == 1) v1 = x; else if (nloc == 2) v2 = x; else v3 = x;
                    nloc++;
                    cur = 0; curk = 0;
                }
            }
        }
        const u32 pk = cur | (curk << 24);
        const u32 E = __ballot_sync(0xffffffffu, nloc > 0u);
        u32 pv, pl;
        sd_incoming(E, lane, pk, carry_x, carry_k, pv, pl);
        if (nloc) {
            v0 = pv | shl32(v0, 5u * pl);
            if (hs) v0 |= shl32(0xffffffffu, 5u * (pl + hk));
        }
        // ---- compaction: number t of this step goes to slot t
        const u32 b0 = __ballot_sync(0xffffffffu, nloc & 1u), b1 = __ballot_sync(0xffffffffu, nloc & 2u),
                  b2 = __ballot_sync(0xffffffffu, nloc & 4u);
        const u32 slot = __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt);
        const u32 nn = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
        if (nloc > 0u) S[slot] = v0;
        if (nloc > 1u) S[slot + 1u] = v1;
        if (nloc > 2u) S[slot + 2u] = v2;
        if (nloc > 3u) S[slot + 3u] = v3;
        __syncwarp();
        const uint4 xs = reinterpret_cast<const uint4 *>(S)[lane];
        __syncwarp();
        // ---- delta: numbers 4 lane .. 4 lane + 3 of the step; slots 0 and 2 are one chain, 1 and 3 the other
        const u32 t0 = 4u * lane, m0 = m_base + t0;
        const u32 x0 = xs.x, x1 = xs.y, x2 = xs.z, x3 = xs.w;
        const u32 y0 = (t0 < nn && m0 != 0u) ? x0 : 0u, y1 = t0 + 1u < nn ? x1 : 0u,
                  y2 = t0 + 2u < nn ? x2 : 0u, y3 = t0 + 3u < nn ? x3 : 0u;
        const u32 A = y0 + y2, B = y1 + y3;
        u32 SA = A, SB = B;
#pragma unroll
        for (u32 d = 1; d < 32; d <<= 1) {
            const u32 ta = __shfl_up_sync(0xffffffffu, SA, d), tb = __shfl_up_sync(0xffffffffu, SB, d);
            if (lane >= d) { SA += ta; SB += tb; }
        }
        const bool odd = m_base & 1u;                            // parity of the numbers in slots 0 and 2
        const u32 cA = odd ? carry_o : carry_e, cB = odd ? carry_e : carry_o;
        const u32 w0 = SA - A + y0 + cA, w2 = w0 + y2, w1 = SB - B + y1 + cB, w3 = w1 + y3;
        if (t0 < nn) out[m0] = m0 == 0u ? x0 : w0;
        if (t0 + 1u < nn) out[m0 + 1u] = w1;
        if (t0 + 2u < nn) out[m0 + 2u] = w2;
        if (t0 + 3u < nn) out[m0 + 3u] = w3;
        m_base += nn;
        if (p0 + 128 < len) {                                    // (uniform) carries into the next step
            const u32 nA = cA + __shfl_sync(0xffffffffu, SA, 31), nB = cB + __shfl_sync(0xffffffffu, SB, 31);
            carry_e = odd ? nB : nA;
            carry_o = odd ? nA : nB;
            sd_incoming(E, 32u, pk, carry_x, carry_k, carry_x, carry_k);
        }
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
    if (n <= SD_SMALL_BATCH)
        rle_string_decode_warp_kernel<<<(unsigned)(((i64)n + SD_WARPS - 1) / SD_WARPS), SD_WARPS * 32, 0, as_stream(stream)>>>(
            d_chars, d_chr_off, n, d_cnt, d_cnt_off, d_cnt_len);
    else
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
