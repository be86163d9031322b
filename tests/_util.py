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


# ---- one-shot forms of the oracle for BASELINE.json's native sizes ------------------------------------------
# oracle/ampis_ref.py restates the reference's loops literally (G x ceil(P/80) RLE.iou calls, S x N merges with a
# full-frame malloc each): minutes to hours at 5,000 x 5,000 or 200 x 2,000 masks of 2048 x 2048.  The functions
# below compute the SAME quantities from ONE call of the oracle's rleIou over all pairs (its bbIou pre-pass included)
# and the oracle's rleMerge on the overlapping pairs only; tests/test_oracle.py pins them to the literal loops on
# small inputs.

def iou_matrix_one_shot(rle, gt, pred):
    """float64[G, P]: rle.iou(pred, gt, iscrowd=0).T in a single call (what analyze.py:108,158 assembles blockwise)."""
    if len(gt) == 0 or len(pred) == 0:
        return np.zeros((len(gt), len(pred)))
    return np.ascontiguousarray(rle.iou(pred, gt, np.zeros(len(gt), np.uint8)).T)


def match_from_iou(iou, thresh):
    """analyze.py:149-179 on a full IoU matrix: per-GT first arg-max, strict '>' against 0 and against thresh."""
    G, P = iou.shape
    if P == 0:
        best, top = np.full(G, -1, np.int64), np.zeros(G)
    else:
        best = iou.argmax(axis=1)
        top = iou[np.arange(G), best] if G else np.zeros(0)
        best = np.where(top > 0, best, -1)
    m = top > thresh
    gi = np.nonzero(m)[0]
    tp = np.stack([gi, best[m]], axis=1).astype(int) if len(gi) else np.asarray([], int)
    pm = np.zeros(P, bool)
    pm[best[m]] = True
    return {'tp': tp, 'fn': np.nonzero(~m)[0].astype(int), 'fp': np.nonzero(~pm)[0].astype(int), 'iou': top[m]}


def det_seg_scores_one_shot(rle, gt, pred, thresh, iou=None):
    """analyze.py:226-339 from the one-shot IoU matrix; merges only for the matched pairs (analyze.py:315)."""
    iou = iou_matrix_one_shot(rle, gt, pred) if iou is None else iou
    det = match_from_iou(iou, thresh)
    tp = np.asarray(det['tp']).reshape(-1, 2)
    n_tp, n_fn, n_fp = len(tp), len(det['fn']), len(det['fp'])
    a_gt, a_pr = rle.area(gt).astype(np.int64), rle.area(pred).astype(np.int64)
    inter = np.array([rle.merge_area(gt[g], pred[p], intersect=True) for g, p in tp], np.int64)
    g_a, p_a = (a_gt[tp[:, 0]], a_pr[tp[:, 1]]) if n_tp else (np.zeros(0, np.int64), np.zeros(0, np.int64))
    with np.errstate(invalid='ignore', divide='ignore'):
        return {'det_precision': n_tp / (n_tp + n_fp), 'det_recall': n_tp / (n_tp + n_fn),
                'seg_precision': inter / (inter + (p_a - inter)), 'seg_recall': inter / (inter + (g_a - inter)),
                'det_tp': det['tp'], 'det_fn': det['fn'], 'det_fp': det['fp'], 'seg_tp': inter,
                'seg_fn': g_a - inter, 'seg_fp': p_a - inter, 'det_tp_iou': det['iou']}


def satellite_match_one_shot(rle, particles, satellites, thresh):
    """powder.py:28-112: intersections only where rleIou is non-zero (disjoint masks intersect in 0 pixels),
    then the reference's per-satellite arg-max over ALL particles (NaN row for an empty satellite)."""
    S, N = len(satellites), len(particles)
    iou = iou_matrix_one_shot(rle, satellites, particles)
    inter = np.zeros((S, N), np.uint32)
    for s, p in zip(*np.nonzero(iou)):
        inter[s, p] = rle.merge_area(satellites[s], particles[p], intersect=True)
    with np.errstate(invalid='ignore', divide='ignore'):
        score = inter / rle.area(satellites).astype(np.uint32)[:, None]
    best = score.argmax(axis=1)            # NaN row: arg-max 0, never above the threshold
    top = score[np.arange(S), best]
    m = top > thresh
    sm = np.stack([np.nonzero(m)[0], best[m]], axis=1).astype(np.int64)
    pm = np.zeros(N, bool)
    pm[best[m]] = True
    pairs = {x: [] for x in np.unique(sm[:, 1])}
    for s_, p_ in sm:
        pairs[p_].append(s_)
    return {'satellite_matches': sm, 'satellites_unmatched': np.nonzero(~m)[0].astype(np.int64),
            'particles_unmatched': np.nonzero(~pm)[0].astype(np.int64), 'intersection_scores': top[m],
            'match_pairs': pairs}
