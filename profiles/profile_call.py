#!/usr/bin/env python
"""Where the per-call time of the drop-in API goes: cProfile of analyze.det_seg_scores on one C1-like
image (300 x 300 masks), 200 warm calls.  python profiles/profile_call.py > gpurun_out/profile_call.txt"""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ampis_b200 import analyze as A, batch as B
    from ampis_b200.applications import powder as P
    from oracle import cocomask as rle
    rle.build()
    torch.cuda.set_device(0)
    host = B.synth(dict(B.CONFIGS['c1_powder_example']), 1, 11)
    r, c = host.image_masks(0)
    size = [host.h, host.w]
    mk = lambda cs: [{'size': size, 'counts': rle.string_from_counts(x)} for x in cs]
    gt, pr = mk(r), mk(c)
    for _ in range(20):
        A.det_seg_scores(gt, pr, 0.5)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        A.det_seg_scores(gt, pr, 0.5)
    torch.cuda.synchronize()
    print('det_seg_scores: %.3f ms per call (%d x %d masks)' % (1e3 * (time.perf_counter() - t0) / n, len(gt), len(pr)))
    t0 = time.perf_counter()
    for _ in range(n):
        A.rle_instance_matcher(gt, pr, 0.5)
    print('rle_instance_matcher: %.3f ms per call' % (1e3 * (time.perf_counter() - t0) / n))
    t0 = time.perf_counter()
    for _ in range(n):
        P._rle_satellite_match(gt, pr[:60], 0.5)
    print('_rle_satellite_match: %.3f ms per call' % (1e3 * (time.perf_counter() - t0) / n))
    pr_ = cProfile.Profile()
    pr_.enable()
    for _ in range(n):
        A.det_seg_scores(gt, pr, 0.5)
    pr_.disable()
    st = pstats.Stats(pr_, stream=sys.stdout)
    st.sort_stats('cumulative').print_stats(45)
    st.sort_stats('tottime').print_stats(30)


if __name__ == '__main__':
    main()
