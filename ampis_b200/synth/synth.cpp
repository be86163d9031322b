// Synthetic micrograph generator (host code; bench / test data only, never on the
// evaluated path).  Emits COCO-order run counts directly, so neither the GPU path nor
// the CPU oracle spends time producing its own input.  See DESIGN.md "Synthetic data".
//
// An image holds n_gt "primary" blobs (powder particles / spheroidite carbides: rotated
// ellipses with a jagged boundary, log-normal equivalent diameters, rejection-placed so
// they rarely overlap) and n_sec "secondary" masks that are either
//   kind 0: predictions of the primaries (centre jitter, radius scale, a fraction dropped
//           and replaced by spurious blobs, order shuffled), or
//   kind 1: satellites (small blobs, most of them straddling the rim of a primary).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/ampis_synth.h"

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next()
    {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    double normal()
    {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    }
};

struct Blob {
    double cx, cy, r, aspect, theta;
    bool empty;
};

// run counts of one blob, appended to out; returns number of counts
int emit_blob(const Blob &b, uint32_t h, uint32_t w, Rng &rng, std::vector<uint32_t> &out)
{
    const size_t first = out.size();
    const uint64_t hw = (uint64_t)h * w;
    uint64_t prev_end = 0;
    bool have_run = false;
    if (!b.empty) {
        const double a = b.r * sqrt(b.aspect), bb = b.r / sqrt(b.aspect);
        const double ct = cos(b.theta), st = sin(b.theta);
        const double A = ct * ct / (a * a) + st * st / (bb * bb);
        const double B = ct * st * (1.0 / (a * a) - 1.0 / (bb * bb));
        const double C = st * st / (a * a) + ct * ct / (bb * bb);
        const double xext = sqrt(C / (A * C - B * B));
        long x0 = (long)ceil(b.cx - xext), x1 = (long)floor(b.cx + xext);
        if (x0 < 0) x0 = 0;
        if (x1 > (long)w - 1) x1 = (long)w - 1;
        for (long x = x0; x <= x1; x++) {
            const double dx = (double)x - b.cx;
            const double D = B * B * dx * dx - C * (A * dx * dx - 1.0);
            if (D < 0) continue;
            const double sq = sqrt(D);
            long ylo = (long)ceil(b.cy + (-B * dx - sq) / C), yhi = (long)floor(b.cy + (-B * dx + sq) / C);
            // jagged boundary: each end moves by one pixel now and then
            const uint64_t rr = rng.next();
            if ((rr & 7) == 0) ylo -= 1; else if ((rr & 7) == 1) ylo += 1;
            if (((rr >> 3) & 7) == 0) yhi += 1; else if (((rr >> 3) & 7) == 1) yhi -= 1;
            if (ylo < 0) ylo = 0;
            if (yhi > (long)h - 1) yhi = (long)h - 1;
            if (yhi < ylo) continue;
            const uint64_t s = (uint64_t)x * h + (uint64_t)ylo, e = (uint64_t)x * h + (uint64_t)yhi + 1;
            if (have_run && s == prev_end) {
                out.back() += (uint32_t)(e - s);   // column ran to the bottom, next starts at the top
            } else {
                out.push_back((uint32_t)(s - prev_end));
                out.push_back((uint32_t)(e - s));
            }
            prev_end = e;
            have_run = true;
        }
    }
    if (!have_run) out.push_back((uint32_t)hw);
    else if (prev_end < hw) out.push_back((uint32_t)(hw - prev_end));
    return (int)(out.size() - first);
}

struct Params {
    uint32_t h, w;
    int n_gt, n_sec, kind;
    double median_diam, sigma_ln, max_aspect, sec_median_diam, jitter_px, scale_sigma, drop_frac, empty_frac;
};

double draw_radius(Rng &rng, double median_diam, double sigma_ln, double lo, double hi)
{
    double d = median_diam * exp(sigma_ln * rng.normal());
    if (d < lo) d = lo;
    if (d > hi) d = hi;
    return 0.5 * d;
}

void gen_image(uint64_t seed, const Params &p, std::vector<uint32_t> &cnt, std::vector<int32_t> &len)
{
    Rng rng(seed * 0x2545f4914f6cdd1dull + 0x1234567ull);
    const double dmax = 0.4 * (double)std::min(p.h, p.w);
    std::vector<Blob> gt(p.n_gt);
    for (auto &b : gt) {
        b.r = draw_radius(rng, p.median_diam, p.sigma_ln, 3.0, dmax);
        b.aspect = 1.0 + (p.max_aspect - 1.0) * rng.uni();
        b.theta = 3.141592653589793 * rng.uni();
        b.empty = rng.uni() < p.empty_frac;
    }
    std::sort(gt.begin(), gt.end(), [](const Blob &a, const Blob &b) { return a.r > b.r; });
    // rejection placement on a uniform grid of cells of size 2*r_max
    const double cell = std::max(2.0 * (gt.empty() ? 1.0 : gt[0].r), 8.0);
    const int gx = (int)(p.w / cell) + 1, gy = (int)(p.h / cell) + 1;
    std::vector<std::vector<int>> grid((size_t)gx * gy);
    for (int i = 0; i < p.n_gt; i++) {
        Blob &b = gt[i];
        for (int tries = 0; tries < 60; tries++) {
            b.cx = rng.uni() * p.w;
            b.cy = rng.uni() * p.h;
            bool ok = true;
            const int cxi = (int)(b.cx / cell), cyi = (int)(b.cy / cell);
            for (int yy = std::max(0, cyi - 1); ok && yy <= std::min(gy - 1, cyi + 1); yy++)
                for (int xx = std::max(0, cxi - 1); ok && xx <= std::min(gx - 1, cxi + 1); xx++)
                    for (int j : grid[(size_t)yy * gx + xx]) {
                        const double dx = gt[j].cx - b.cx, dy = gt[j].cy - b.cy, rr = 0.93 * (gt[j].r + b.r);
                        if (dx * dx + dy * dy < rr * rr) { ok = false; break; }
                    }
            if (ok) break;
        }
        grid[(size_t)((int)(b.cy / cell)) * gx + (int)(b.cx / cell)].push_back(i);
    }
    // primaries are listed in a shuffled order (not by size)
    for (int i = p.n_gt - 1; i > 0; i--) std::swap(gt[i], gt[rng.next() % (uint64_t)(i + 1)]);

    std::vector<Blob> sec;
    sec.reserve(p.n_sec);
    if (p.kind == 0) {
        for (int i = 0; i < p.n_gt && (int)sec.size() < p.n_sec; i++) {
            if (rng.uni() < p.drop_frac) continue;
            Blob b = gt[i];
            b.cx += p.jitter_px * rng.normal();
            b.cy += p.jitter_px * rng.normal();
            double sc = 1.0 + p.scale_sigma * rng.normal();
            b.r *= std::min(1.5, std::max(0.5, sc));
            b.theta += 0.05 * rng.normal();
            b.empty = rng.uni() < p.empty_frac;
            sec.push_back(b);
        }
        while ((int)sec.size() < p.n_sec) {   // spurious detections
            Blob b;
            b.r = draw_radius(rng, p.median_diam, p.sigma_ln, 3.0, dmax);
            b.aspect = 1.0 + (p.max_aspect - 1.0) * rng.uni();
            b.theta = 3.141592653589793 * rng.uni();
            b.cx = rng.uni() * p.w;
            b.cy = rng.uni() * p.h;
            b.empty = rng.uni() < p.empty_frac;
            sec.push_back(b);
        }
    } else {
        for (int i = 0; i < p.n_sec; i++) {
            Blob b;
            b.r = draw_radius(rng, p.sec_median_diam, 0.5, 2.0, dmax);
            b.aspect = 1.0 + 0.3 * rng.uni();
            b.theta = 3.141592653589793 * rng.uni();
            b.empty = rng.uni() < p.empty_frac;
            if (p.n_gt > 0 && rng.uni() < 0.85) {
                const Blob &host = gt[rng.next() % (uint64_t)p.n_gt];
                const double ang = 6.283185307179586 * rng.uni(), rad = host.r * (0.75 + 0.35 * rng.uni());
                b.cx = host.cx + rad * cos(ang);
                b.cy = host.cy + rad * sin(ang);
            } else {
                b.cx = rng.uni() * p.w;
                b.cy = rng.uni() * p.h;
            }
            sec.push_back(b);
        }
    }
    for (int i = (int)sec.size() - 1; i > 0; i--) std::swap(sec[i], sec[rng.next() % (uint64_t)(i + 1)]);

    len.clear();
    for (const Blob &b : gt) len.push_back(emit_blob(b, p.h, p.w, rng, cnt));
    for (const Blob &b : sec) len.push_back(emit_blob(b, p.h, p.w, rng, cnt));
}

}  // namespace

extern "C" int64_t ampis_synth_batch(uint64_t seed, int32_t n_images, uint32_t h, uint32_t w, int32_t n_gt,
                                     int32_t n_sec, int32_t kind, double median_diam, double sigma_ln,
                                     double max_aspect, double sec_median_diam, double jitter_px,
                                     double scale_sigma, double drop_frac, double empty_frac, int32_t n_threads,
                                     uint32_t *cnt, int64_t cnt_capacity, int64_t *cnt_off, int32_t *cnt_len)
{
    if (n_images < 0 || n_gt < 0 || n_sec < 0 || h == 0 || w == 0 || (kind != 0 && kind != 1)) return 0;      // nothing to generate
    Params p{h, w, n_gt, n_sec, kind, median_diam, sigma_ln, max_aspect, sec_median_diam,
             jitter_px, scale_sigma, drop_frac, empty_frac};
    std::vector<std::vector<uint32_t>> cnts(n_images);
    std::vector<std::vector<int32_t>> lens(n_images);
    if (n_threads < 1) n_threads = 1;
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++)
        pool.emplace_back([&, t]() {
            for (int i = t; i < n_images; i += n_threads) gen_image(seed + (uint64_t)i, p, cnts[i], lens[i]);
        });
    for (auto &th : pool) th.join();
    int64_t total = 0;
    for (auto &c : cnts) total += (int64_t)c.size();
    if (total > cnt_capacity) return -total;   // caller retries with |value| capacity
    int64_t off = 0, k = 0;
    const int per = n_gt + n_sec;
    for (int i = 0; i < n_images; i++) {
        memcpy(cnt + off, cnts[i].data(), cnts[i].size() * sizeof(uint32_t));
        int64_t o = off;
        for (int j = 0; j < per; j++) {
            cnt_off[k] = o;
            cnt_len[k] = lens[i][j];
            o += lens[i][j];
            k++;
        }
        off += (int64_t)cnts[i].size();
    }
    return total;
}
