"""Shared helpers for the test-suite: golden fixture access and a dense numpy
formulation (bit-packed AND + popcount) that is independent of the run-walk."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def unpack_strings(blob, off, size):
    b = blob.tobytes()
    return [{'size': [int(size[0]), int(size[1])], 'counts': b[off[i]:off[i + 1]]} for i in range(len(off) - 1)]


def powder_match_image(k):
    g = load('powder_match.npz')
    size = g['%d_size' % k]
    gt = unpack_strings(g['%d_gt_blob' % k], g['%d_gt_off' % k], size)
    pr = unpack_strings(g['%d_pr_blob' % k], g['%d_pr_off' % k], size)
    return g, gt, pr


def powder_satellite_image(k):
    s = load('powder_satellite.npz')
    g = load('powder_match.npz')
    names = list(g['names'])
    kk = names.index(str(s['%d_part_ref' % k]))
    size = s['%d_size' % k]
    part = unpack_strings(g['%d_pr_blob' % kk], g['%d_pr_off' % kk], size)
    sat = unpack_strings(s['%d_sat_blob' % k], s['%d_sat_off' % k], size)
    return s, part, sat


def counts_to_packed(cnts, hw):
    """uint32 counts -> np.packbits (little bit order) of the column-major bit vector"""
    ends = np.cumsum(cnts.astype(np.int64))
    bits = np.zeros(hw + 1, np.int8)
    starts = ends[:-1]
    np.add.at(bits, starts[starts <= hw], 1)
    v = (np.cumsum(bits[:hw]) & 1).astype(np.uint8)
    return np.packbits(v, bitorder='little')


def popcount(a):
    return int(np.bitwise_count(a).sum())


def rand_masks(rng, n, h, w, p_empty=0.1):
    """n random blob-ish bool masks [n,h,w]"""
    out = np.zeros((n, h, w), bool)
    yy, xx = np.mgrid[0:h, 0:w]
    for i in range(n):
        if rng.random() < p_empty:
            continue
        for _ in range(rng.integers(1, 4)):
            cy, cx = rng.uniform(0, h), rng.uniform(0, w)
            ry, rx = rng.uniform(0.5, h / 2), rng.uniform(0.5, w / 2)
            out[i] |= ((yy - cy) / ry) ** 2 + ((xx - cx) / rx) ** 2 <= 1
        if rng.random() < 0.3:
            out[i] ^= rng.random((h, w)) < 0.05
    return out
