/*
 * ampis_synth.h -- synthetic micrograph generator (libampis_synth.so): BENCH / TEST DATA ONLY.
 * Plain host C++, no CUDA, deliberately a separate library: bench.py's reference arm and the CPU tests generate
 * their inputs without ever mapping libampis_b200.so (the product).
 */
#ifndef AMPIS_SYNTH_H
#define AMPIS_SYNTH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Deterministic powder-like images: per image n_gt primary blobs followed by n_sec secondary
 * masks (kind 0: predictions of the primaries -- jitter, scale, drop_frac dropped and replaced
 * by spurious blobs, shuffled; kind 1: satellites, 85 % on primary rims).  Writes the run
 * counts of all n_images*(n_gt+n_sec) masks (image-major, primaries first) as CSR.
 * Returns the number of counts written; if cnt_capacity is too small returns -(needed). */
int64_t ampis_synth_batch(uint64_t seed, int32_t n_images, uint32_t h, uint32_t w, int32_t n_gt,
                          int32_t n_sec, int32_t kind, double median_diam, double sigma_ln,
                          double max_aspect, double sec_median_diam, double jitter_px,
                          double scale_sigma, double drop_frac, double empty_frac, int32_t n_threads,
                          uint32_t *cnt, int64_t cnt_capacity, int64_t *cnt_off, int32_t *cnt_len);

#ifdef __cplusplus
}
#endif
#endif /* AMPIS_SYNTH_H */
