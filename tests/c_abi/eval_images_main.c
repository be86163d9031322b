/* C caller of libampis_b200.so without Python or PyTorch, MANY images per call: reads a dataset (compressed RLE
 * strings back to back, per-image counts and sizes) from a binary file, runs ampis_eval_images_host once with
 * cudaMalloc'd / cudaHostAlloc'd workspaces -- strings passed as ONE contiguous blob (AMPIS_STRINGS_CONTIGUOUS), and,
 * with a second argument, as one (pointer, length) descriptor per string -- and prints the per-row results and the
 * TP / FP / FN counts at two thresholds.  Built and run by tests/test_gpu_parity.py::test_c_caller_of_the_abi.
 *   file: int32 n_images; int32 n_rows[n_images], n_cols[n_images], h[n_images], w[n_images];
 *         int32 str_len[n]; bytes chars[sum(str_len)]                                                         */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <cuda_runtime.h>
#include "ampis_b200.h"

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    const int scattered = argc > 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t ni;
    if (fread(&ni, 4, 1, f) != 1) return 2;
    int32_t *n_rows = malloc(4 * (size_t)ni), *n_cols = malloc(4 * (size_t)ni);
    uint32_t *h = malloc(4 * (size_t)ni), *w = malloc(4 * (size_t)ni);
    if (fread(n_rows, 4, ni, f) != (size_t)ni || fread(n_cols, 4, ni, f) != (size_t)ni ||
        fread(h, 4, ni, f) != (size_t)ni || fread(w, 4, ni, f) != (size_t)ni) return 2;
    int64_t n = 0, R = 0, n_chars = 0;
    for (int32_t g = 0; g < ni; g++) { n += n_rows[g] + n_cols[g]; R += n_rows[g]; }
    int32_t *len = malloc(4 * (size_t)n);
    if (fread(len, 4, (size_t)n, f) != (size_t)n) return 2;
    for (int64_t i = 0; i < n; i++) n_chars += len[i];
    uint8_t *chars = NULL;
    if (cudaHostAlloc((void **)&chars, (size_t)n_chars + 1, cudaHostAllocDefault) != cudaSuccess) return 3;
    if (fread(chars, 1, (size_t)n_chars, f) != (size_t)n_chars) return 2;
    fclose(f);
    const uint8_t **ptr = malloc(8 * (size_t)n);
    { int64_t pos = 0; for (int64_t i = 0; i < n; i++) { ptr[i] = chars + pos; pos += len[i]; } }

    const double thr[2] = {0.5, 0.75};
    int32_t *best_col = malloc(4 * (size_t)(R + 1)), *status = malloc(4 * (size_t)n), *counts = malloc(12 * 2 * (size_t)ni);
    uint32_t *best_inter = malloc(4 * (size_t)(R + 1)), *area = malloc(4 * (size_t)n);
    double *best_score = malloc(8 * (size_t)(R + 1));
    int64_t totals[6], found = 0, need = 0, d_bytes = 1 << 20, h_bytes = 1 << 16;
    int32_t crowded = 0;
    void *d_ws = NULL, *h_ws = NULL;
    int rc = AMPIS_ENOSPC;
    for (int attempt = 0; attempt < 8 && rc == AMPIS_ENOSPC; attempt++) {
        if (need < 0) h_bytes = -need * 2;
        if (need > 0) d_bytes = need * 2;
        if (d_ws) cudaFree(d_ws);
        if (h_ws) cudaFreeHost(h_ws);
        if (cudaMalloc(&d_ws, (size_t)d_bytes) != cudaSuccess) return 3;
        if (cudaHostAlloc(&h_ws, (size_t)h_bytes, cudaHostAllocDefault) != cudaSuccess) return 3;
        rc = ampis_eval_images_host(ptr, len, ni, n_rows, n_cols, h, w, AMPIS_MODE_IOU,
                                    scattered ? 0 : AMPIS_STRINGS_CONTIGUOUS, -1.0, d_ws, d_bytes, h_ws, h_bytes,
                                    best_col, best_inter, best_score, area, NULL, NULL, status, thr, 2, counts, totals,
                                    &found, &crowded, &need, NULL /* default stream */);
    }
    if (rc != AMPIS_OK) { fprintf(stderr, "ampis_eval_images_host: %d %s\n", rc, ampis_last_error()); return 1; }
    for (int64_t r = 0; r < R; r++) printf("row %lld %d %u %.17g\n", (long long)r, best_col[r], best_inter[r], best_score[r]);
    for (int64_t i = 0; i < n; i++) printf("mask %lld %u %d\n", (long long)i, area[i], status[i]);
    for (int32_t g = 0; g < ni; g++)
        printf("counts %d %d %d %d %d %d %d\n", g, counts[6 * g], counts[6 * g + 1], counts[6 * g + 2], counts[6 * g + 3],
               counts[6 * g + 4], counts[6 * g + 5]);
    printf("totals %lld %lld %lld %lld %lld %lld pairs %lld crowded %d\n", (long long)totals[0], (long long)totals[1],
           (long long)totals[2], (long long)totals[3], (long long)totals[4], (long long)totals[5], (long long)found, crowded);
    cudaFree(d_ws);
    cudaFreeHost(h_ws);
    cudaFreeHost(chars);
    return 0;
}
