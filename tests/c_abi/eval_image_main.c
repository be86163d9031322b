/* C caller of libampis_b200.so without Python or PyTorch: reads one image (compressed RLE strings of n_rows +
 * n_cols masks) from a binary file, runs ampis_eval_image_host with cudaMalloc'd / cudaHostAlloc'd workspaces and
 * prints the per-row results.  Built and run by tests/test_gpu_parity.py::test_c_caller_of_the_abi.
 *   file: int32 n_rows, n_cols, h, w, mode; int64 chr_off[n+1]; bytes chars[chr_off[n]]                        */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <cuda_runtime.h>
#include "ampis_b200.h"

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t hdr[5];
    if (fread(hdr, 4, 5, f) != 5) return 2;
    const int32_t n_rows = hdr[0], n_cols = hdr[1], n = n_rows + n_cols;
    int64_t *off = (int64_t *)malloc(8 * (size_t)(n + 1));
    if (fread(off, 8, (size_t)n + 1, f) != (size_t)n + 1) return 2;
    uint8_t *chars = (uint8_t *)malloc((size_t)off[n] + 1);
    if (fread(chars, 1, (size_t)off[n], f) != (size_t)off[n]) return 2;
    fclose(f);

    int64_t d_bytes = 1 << 20, h_bytes = 1 << 16, need = 0;
    void *d_ws = NULL, *h_ws = NULL;
    int32_t *best_col = (int32_t *)malloc(4 * (size_t)(n_rows + 1));
    uint32_t *best_inter = (uint32_t *)malloc(4 * (size_t)(n_rows + 1));
    double *best_score = (double *)malloc(8 * (size_t)(n_rows + 1));
    uint32_t *area = (uint32_t *)malloc(4 * (size_t)n), *span = (uint32_t *)malloc(8 * (size_t)n);
    int32_t *bbox = (int32_t *)malloc(16 * (size_t)n), *status = (int32_t *)malloc(4 * (size_t)n);
    int rc = AMPIS_ENOSPC;
    for (int attempt = 0; attempt < 8 && rc == AMPIS_ENOSPC; attempt++) {
        if (need < 0) h_bytes = -need * 2;
        if (need > 0) d_bytes = need * 2;
        if (d_ws) cudaFree(d_ws);
        if (h_ws) cudaFreeHost(h_ws);
        if (cudaMalloc(&d_ws, (size_t)d_bytes) != cudaSuccess) return 3;
        if (cudaHostAlloc(&h_ws, (size_t)h_bytes, cudaHostAllocDefault) != cudaSuccess) return 3;
        rc = ampis_eval_image_host(chars, off, n_rows, n_cols, (uint32_t)hdr[2], (uint32_t)hdr[3], hdr[4], 64,
                                   d_ws, d_bytes, h_ws, h_bytes, best_col, best_inter, best_score, area, bbox, span,
                                   status, NULL, &need, NULL /* default stream */);
    }
    if (rc != AMPIS_OK) { fprintf(stderr, "ampis_eval_image_host: %d %s\n", rc, ampis_last_error()); return 1; }
    for (int32_t r = 0; r < n_rows; r++) printf("row %d %d %u %.17g\n", r, best_col[r], best_inter[r], best_score[r]);
    for (int32_t i = 0; i < n; i++)
        printf("mask %d %u %d %d %d %d %d\n", i, area[i], bbox[4 * i], bbox[4 * i + 1], bbox[4 * i + 2], bbox[4 * i + 3], status[i]);
    cudaFree(d_ws);
    cudaFreeHost(h_ws);
    return 0;
}
