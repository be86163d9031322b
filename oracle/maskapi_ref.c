/*
 * oracle/maskapi_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (C99, scalar, single thread) of the run-length mask
 * arithmetic that AMPIS delegates to its third-party dependency
 * pycocotools==2.0.4 (pinned at /root/reference/docker/env.yml:21; files
 * common/maskApi.c + pycocotools/_mask.pyx, NOT vendored under
 * /root/reference).  The published algorithm is restated here from its
 * documented behaviour; AMPIS's own call sites are
 * analyze.py:108,158,315,320,321; powder.py:82,83,264;
 * structures.py:465,467,568,571,677,752,761; data_utils.py:275,423,514.
 *
 * PARITY STATUS: "parity unpinned" against a real pycocotools binary (none
 * is installable here: no wheel, no network).  What IS pinned: the only
 * known-answer test in the reference (analyze.py:702-728), structural
 * invariants on all 6,012 fixture RLE strings, and a second independent
 * formulation (dense numpy bitmasks) -- see tests/test_oracle_*.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Conventions: a mask is h*w pixels in column-major order (index = x*h+y);
 * cnts[0] counts leading zeros (may be 0), runs then alternate 1,0,1,...
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint32_t u32;
typedef unsigned long siz;

static u32 u32min(u32 a, u32 b) { return a < b ? a : b; }
static u32 u32max(u32 a, u32 b) { return a > b ? a : b; }

/* ---- A.2 compressed string codec ------------------------------------- */

/* rleFrString: 6-bit chars (ascii 48..111), 5 payload bits + continuation
 * bit 0x20, sign bit 0x10 of the last group; counts with index > 2 are
 * stored as a delta against cnts[i-2].  Returns m. out needs len slots. */
siz ref_rle_fr_string(const char *s, siz len, u32 *out)
{
    siz m = 0, p = 0;
    while (p < len && s[p]) {
        long x = 0;
        int k = 0, more = 1;
        while (more) {
            long c = (long)s[p] - 48;
            x |= (c & 0x1f) << (5 * k);
            more = (int)(c & 0x20);
            p++;
            k++;
            if (!more && (c & 0x10))
                x |= (long)(~0UL << (5 * k));
        }
        if (m > 2)
            x += (long)out[m - 2];
        out[m++] = (u32)x;
    }
    return m;
}

/* rleToString: inverse of the above. out needs 6*m+1 bytes. Returns strlen. */
siz ref_rle_to_string(const u32 *cnts, siz m, char *out)
{
    siz p = 0;
    for (siz i = 0; i < m; i++) {
        long x = (long)cnts[i];
        if (i > 2)
            x -= (long)cnts[i - 2];
        int more = 1;
        while (more) {
            char c = (char)(x & 0x1f);
            x >>= 5;
            more = (c & 0x10) ? x != -1 : x != 0;
            if (more)
                c |= 0x20;
            out[p++] = (char)(c + 48);
        }
    }
    out[p] = 0;
    return p;
}

/* ---- A.3 area / decode / encode -------------------------------------- */

u32 ref_rle_area(const u32 *cnts, siz m)
{
    u32 a = 0;
    for (siz j = 1; j < m; j += 2)
        a += cnts[j];
    return a;
}

/* writes sum(cnts) bytes (column-major), value = run parity */
void ref_rle_decode(const u32 *cnts, siz m, uint8_t *M)
{
    uint8_t v = 0;
    for (siz j = 0; j < m; j++) {
        for (u32 k = 0; k < cnts[j]; k++)
            *(M++) = v;
        v = !v;
    }
}

/* one mask of h*w column-major bytes -> counts; out needs h*w+1 slots */
siz ref_rle_encode(const uint8_t *M, siz h, siz w, u32 *out)
{
    siz a = h * w, k = 0;
    u32 c = 0;
    uint8_t p = 0;
    for (siz j = 0; j < a; j++) {
        if (M[j] != p) {
            out[k++] = c;
            c = 0;
            p = M[j];
        }
        c++;
    }
    out[k++] = c;
    return k;
}

/* ---- A.5 merge of two masks (the only arity AMPIS uses) -------------- */

/* Two-pointer walk; like the original it sizes its scratch for the worst
 * case (h*w+1 counts) on every call -- that malloc is part of the CPU cost
 * the reference pays.  out needs mA+mB+1 slots.
 * Returns m of the merged RLE. */
siz ref_rle_merge2(const u32 *A, siz mA, const u32 *B, siz mB, siz h, siz w,
                   int intersect, u32 *out)
{
    u32 *scratch = (u32 *)malloc(sizeof(u32) * (h * w + 1));
    u32 ca = A[0], cb = B[0], cc = 0, ct = 1;
    int v = 0, va = 0, vb = 0, vp;
    siz a = 1, b = 1, m = 0;
    while (ct > 0) {
        u32 c = u32min(ca, cb);
        cc += c;
        ct = 0;
        ca -= c;
        if (!ca && a < mA) { ca = A[a++]; va = !va; }
        ct += ca;
        cb -= c;
        if (!cb && b < mB) { cb = B[b++]; vb = !vb; }
        ct += cb;
        vp = v;
        v = intersect ? (va && vb) : (va || vb);
        if (v != vp || ct == 0) { scratch[m++] = cc; cc = 0; }
    }
    memcpy(out, scratch, sizeof(u32) * m);
    free(scratch);
    return m;
}

/* ---- A.4 bbox, bbIou, rleIou ------------------------------------------ */

void ref_rle_to_bbox(const u32 *cnts, siz m_in, u32 h, u32 w, double *bb)
{
    u32 xs = w, ys = h, xe = 0, ye = 0, xp = 0, cc = 0;
    siz m = (m_in / 2) * 2;
    if (m == 0) { bb[0] = bb[1] = bb[2] = bb[3] = 0; return; }
    for (siz j = 0; j < m; j++) {
        cc += cnts[j];
        u32 t = cc - (u32)(j % 2);
        u32 y = t % h, x = (t - y) / h;
        if (j % 2 == 0)
            xp = x;
        else if (xp < x) { ys = 0; ye = h - 1; }
        xs = u32min(xs, x); xe = u32max(xe, x);
        ys = u32min(ys, y); ye = u32max(ye, y);
    }
    bb[0] = xs; bb[2] = xe - xs + 1;
    bb[1] = ys; bb[3] = ye - ys + 1;
}

/* dt: m boxes, gt: n boxes, (x,y,w,h) doubles; o[g*m+d] */
void ref_bb_iou(const double *dt, const double *gt, siz m, siz n,
                const uint8_t *iscrowd, double *o)
{
    for (siz g = 0; g < n; g++) {
        const double *G = gt + 4 * g;
        double ga = G[2] * G[3];
        int crowd = iscrowd != NULL && iscrowd[g];
        for (siz d = 0; d < m; d++) {
            const double *D = dt + 4 * d;
            double da = D[2] * D[3];
            o[g * m + d] = 0;
            double ww = fmin(D[2] + D[0], G[2] + G[0]) - fmax(D[0], G[0]);
            if (ww <= 0) continue;
            double hh = fmin(D[3] + D[1], G[3] + G[1]) - fmax(D[1], G[1]);
            if (hh <= 0) continue;
            double i = ww * hh;
            double u = crowd ? da : da + ga - i;
            o[g * m + d] = i / u;
        }
    }
}

/* Arrays of RLEs given as pointer/length tables. All masks share (h,w) per
 * entry via hs/ws. o[g*m+d] as in the original. */
void ref_rle_iou(const u32 *const *dt, const siz *dtm, const u32 *dth, const u32 *dtw, siz m,
                 const u32 *const *gt, const siz *gtm, const u32 *gth, const u32 *gtw, siz n,
                 const uint8_t *iscrowd, double *o)
{
    double *db = (double *)malloc(sizeof(double) * 4 * (m ? m : 1));
    double *gb = (double *)malloc(sizeof(double) * 4 * (n ? n : 1));
    for (siz d = 0; d < m; d++) ref_rle_to_bbox(dt[d], dtm[d], dth[d], dtw[d], db + 4 * d);
    for (siz g = 0; g < n; g++) ref_rle_to_bbox(gt[g], gtm[g], gth[g], gtw[g], gb + 4 * g);
    ref_bb_iou(db, gb, m, n, iscrowd, o);
    free(db);
    free(gb);
    for (siz g = 0; g < n; g++)
        for (siz d = 0; d < m; d++) {
            if (!(o[g * m + d] > 0)) continue;
            int crowd = iscrowd != NULL && iscrowd[g];
            if (dth[d] != gth[g] || dtw[d] != gtw[g]) { o[g * m + d] = -1; continue; }
            const u32 *A = dt[d], *B = gt[g];
            siz ka = dtm[d], kb = gtm[g], a = 1, b = 1;
            u32 ca = A[0], cb = B[0], ct = 1, i = 0, u = 0;
            int va = 0, vb = 0;
            while (ct > 0) {
                u32 c = u32min(ca, cb);
                if (va || vb) { u += c; if (va && vb) i += c; }
                ct = 0;
                ca -= c;
                if (!ca && a < ka) { ca = A[a++]; va = !va; }
                ct += ca;
                cb -= c;
                if (!cb && b < kb) { cb = B[b++]; vb = !vb; }
                ct += cb;
            }
            if (i == 0) u = 1;
            else if (crowd) u = ref_rle_area(A, ka);
            o[g * m + d] = (double)i / (double)u;
        }
}

/* ---- A.7 polygon rasteriser ------------------------------------------- */

static int cmp_u32(const void *a, const void *b)
{
    u32 c = *(const u32 *)a, d = *(const u32 *)b;
    return c > d ? 1 : c < d ? -1 : 0;
}

/* (int) of a double the way x86-64 cvttsd2si does it, made explicit so the
 * degenerate zero-length edge (0/0 slope) is deterministic everywhere. */
static int trunc_to_int(double v)
{
    if (!(v > -2147483649.0 && v < 2147483648.0)) return (int)0x80000000;
    return (int)v;
}

/* xy: k vertices (x0,y0,x1,y1,...). out needs cap slots; returns m, or
 * (siz)-1 if cap is too small. */
siz ref_rle_fr_poly(const double *xy, siz k, siz h, siz w, u32 *out, siz cap)
{
    const double scale = 5;
    siz j, m = 0;
    int *x = (int *)malloc(sizeof(int) * (k + 1));
    int *y = (int *)malloc(sizeof(int) * (k + 1));
    for (j = 0; j < k; j++) x[j] = trunc_to_int(scale * xy[2 * j + 0] + .5);
    x[k] = x[0];
    for (j = 0; j < k; j++) y[j] = trunc_to_int(scale * xy[2 * j + 1] + .5);
    y[k] = y[0];
    for (j = 0; j < k; j++) {
        int adx = abs(x[j] - x[j + 1]), ady = abs(y[j] - y[j + 1]);
        m += (siz)(adx > ady ? adx : ady) + 1;
    }
    int *u = (int *)malloc(sizeof(int) * (m ? m : 1));
    int *v = (int *)malloc(sizeof(int) * (m ? m : 1));
    m = 0;
    for (j = 0; j < k; j++) {
        int xs = x[j], xe = x[j + 1], ys = y[j], ye = y[j + 1], t, d;
        int dx = abs(xe - xs), dy = abs(ys - ye);
        int flip = (dx >= dy && xs > xe) || (dx < dy && ys > ye);
        if (flip) { t = xs; xs = xe; xe = t; t = ys; ys = ye; ye = t; }
        double s = dx >= dy ? (double)(ye - ys) / dx : (double)(xe - xs) / dy;
        if (dx >= dy)
            for (d = 0; d <= dx; d++) {
                t = flip ? dx - d : d;
                u[m] = t + xs;
                v[m] = trunc_to_int(ys + s * t + .5);
                m++;
            }
        else
            for (d = 0; d <= dy; d++) {
                t = flip ? dy - d : d;
                v[m] = t + ys;
                u[m] = trunc_to_int(xs + s * t + .5);
                m++;
            }
    }
    free(x);
    free(y);
    /* crossings of the up-sampled boundary with pixel-column centres */
    siz npts = m;
    u32 *a = (u32 *)malloc(sizeof(u32) * (npts + 1));
    m = 0;
    for (j = 1; j < npts; j++) {
        if (u[j] == u[j - 1]) continue;
        double xd = (double)(u[j] < u[j - 1] ? u[j] : u[j] - 1);
        xd = (xd + .5) / scale - .5;
        if (floor(xd) != xd || xd < 0 || xd > (double)w - 1) continue;
        double yd = (double)(v[j] < v[j - 1] ? v[j] : v[j - 1]);
        yd = (yd + .5) / scale - .5;
        if (yd < 0) yd = 0;
        else if (yd > (double)h) yd = (double)h;
        yd = ceil(yd);
        a[m++] = (u32)((int)xd * (int)h + (int)yd);
    }
    free(u);
    free(v);
    siz n = m;
    a[n++] = (u32)(h * w);
    qsort(a, n, sizeof(u32), cmp_u32);
    u32 p = 0;
    for (j = 0; j < n; j++) { u32 t = a[j]; a[j] -= p; p = t; }
    /* drop zero-length runs by fusing their neighbours */
    u32 *b = (u32 *)malloc(sizeof(u32) * n);
    j = 0; m = 0;
    b[m++] = a[j++];
    while (j < n) {
        if (a[j] > 0) b[m++] = a[j++];
        else { j++; if (j < n) b[m - 1] += a[j++]; }
    }
    siz ret = m;
    if (m <= cap) memcpy(out, b, sizeof(u32) * m);
    else ret = (siz)-1;
    free(a);
    free(b);
    return ret;
}

/* rleFrBbox: (x,y,w,h) box -> 4-vertex polygon -> rleFrPoly */
siz ref_rle_fr_bbox(const double *bb, siz h, siz w, u32 *out, siz cap)
{
    double xs = bb[0], xe = xs + bb[2], ys = bb[1], ye = ys + bb[3];
    double xy[8] = { xs, ys, xs, ye, xe, ye, xe, ys };
    return ref_rle_fr_poly(xy, 4, h, w, out, cap);
}

/* ---- _mask.pyx-level entry points (one C call per Python call) -------- */

/* iou(dt, gt, iscrowd) on compressed strings: parses every string on every
 * call exactly as _mask.pyx's _frString does, then rleIou.  All masks are
 * (h,w).  Strings are NUL-terminated.  o has m*n doubles, o[g*m+d]. */
void ref_iou_strings(const char *const *dts, siz m, const char *const *gts, siz n,
                     u32 h, u32 w, const uint8_t *iscrowd, double *o)
{
    siz tot = m + n;
    u32 **c = (u32 **)calloc(tot ? tot : 1, sizeof(u32 *));
    siz *len = (siz *)calloc(tot ? tot : 1, sizeof(siz));
    u32 *hs = (u32 *)calloc(tot ? tot : 1, sizeof(u32));
    u32 *ws = (u32 *)calloc(tot ? tot : 1, sizeof(u32));
    for (siz i = 0; i < tot; i++) {
        const char *s = i < m ? dts[i] : gts[i - m];
        siz L = strlen(s);
        c[i] = (u32 *)malloc(sizeof(u32) * (L ? L : 1));
        len[i] = ref_rle_fr_string(s, L, c[i]);
        hs[i] = h;
        ws[i] = w;
    }
    ref_rle_iou((const u32 *const *)c, len, hs, ws, m,
                (const u32 *const *)(c + m), len + m, hs + m, ws + m, n, iscrowd, o);
    for (siz i = 0; i < tot; i++) free(c[i]);
    free(c); free(len); free(hs); free(ws);
}

/* area(merge([a, b], intersect)) on compressed strings, with the same
 * string round trip the Python wrapper performs:
 * frString x2 -> rleMerge -> toString -> (returned dict) -> frString -> rleArea */
u32 ref_merge2_area_strings(const char *sa, const char *sb, siz h, siz w, int intersect)
{
    siz La = strlen(sa), Lb = strlen(sb);
    u32 *A = (u32 *)malloc(sizeof(u32) * (La ? La : 1));
    u32 *B = (u32 *)malloc(sizeof(u32) * (Lb ? Lb : 1));
    siz mA = ref_rle_fr_string(sa, La, A), mB = ref_rle_fr_string(sb, Lb, B);
    u32 *M = (u32 *)malloc(sizeof(u32) * (mA + mB + 1));
    siz m = ref_rle_merge2(A, mA, B, mB, h, w, intersect, M);
    char *s = (char *)malloc(6 * m + 1);
    siz L = ref_rle_to_string(M, m, s);
    u32 *R = (u32 *)malloc(sizeof(u32) * (L ? L : 1));
    siz mr = ref_rle_fr_string(s, L, R);
    u32 area = ref_rle_area(R, mr);
    free(A); free(B); free(M); free(s); free(R);
    return area;
}

/* area(list of compressed strings) */
void ref_area_strings(const char *const *ss, siz n, u32 *out)
{
    for (siz i = 0; i < n; i++) {
        siz L = strlen(ss[i]);
        u32 *c = (u32 *)malloc(sizeof(u32) * (L ? L : 1));
        siz m = ref_rle_fr_string(ss[i], L, c);
        out[i] = ref_rle_area(c, m);
        free(c);
    }
}
