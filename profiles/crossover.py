#!/usr/bin/env python
"""Culled AND+popc rows kernels vs dense tcgen05 contraction vs TMA-staged tiled AND+popc as a function of crowding.

    python profiles/crossover.py > gpurun_out/crossover.json      (on the B200)

For a fixed 256x256 frame with 512 x 512 instances the median instance diameter is swept; for
every point both intersection kernels run on the same painted masks (span layout) and are timed
with CUDA events (3 warm-up + 10 timed launches).  `fill` is the fraction of the dense operand
volume that holds mask pixels (sum of span lengths / (masks x slabs)), the quantity
engine.choose_kernel() thresholds; `cand` is the fraction of pairs whose boxes overlap."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from ampis_b200 import batch, engine
    dev = torch.device('cuda', 0)
    out = []
    for frame, n, diam in [(256, 512, 12), (256, 512, 24), (256, 512, 40), (256, 512, 64), (256, 512, 96),
                           (256, 512, 128), (512, 512, 64), (512, 512, 128), (1024, 500, 34), (1024, 500, 128)]:
        cfg = dict(batch.CONFIGS['dense_overlap'], h=frame, w=frame, n_rows=n, n_cols=n, median_diam=float(diam))
        n_img = 64 if frame <= 512 else 37
        host = batch.synth(cfg, n_img, 99)
        db = batch.DeviceBatch(host, dev, dense=True)
        t = engine.MaskTable(dev, host.n_masks, db.cnt, db.cnt_off, db.cnt_len, db.h, db.w, engine.LAYOUT_SPAN)
        t.measure().paint().check()
        fill = engine.operand_fill(t, db.groups)
        bb = t.bbox[:4 * host.n_masks].view(n_img, 2 * n, 4).cpu().numpy().astype(np.int64)
        r, c = bb[:, :n, None, :], bb[:, None, n:, :]
        cand = float(((np.maximum(r[..., 0], c[..., 0]) <= np.minimum(r[..., 2], c[..., 2])) &
                      (np.maximum(r[..., 1], c[..., 1]) <= np.minimum(r[..., 3], c[..., 3]))).mean())
        res = {}
        ref = None
        # the TMA-staged shared-memory tiled AND+popc kernel reads regular full frames: its own table
        tf = engine.MaskTable(dev, host.n_masks, db.cnt, db.cnt_off, db.cnt_len, db.h, db.w, engine.LAYOUT_FULL)
        tf.measure().paint().check()
        tc = engine.MaskTable(dev, host.n_masks, db.cnt, db.cnt_off, db.cnt_len, db.h, db.w, engine.LAYOUT_CROP)
        tc.measure().paint().check()
        grid = engine.ColumnGrid(dev, db.groups.n_groups).build(tc, db.groups)
        pairs_list = engine.intersect_rows(tc, db.groups, db.mode, grid=grid).pairs
        runs = (('rows', lambda out=None: engine.intersect_rows(t, db.groups, db.mode, out=out)),
                ('mma', lambda out=None: engine.intersect_mma(t, db.groups, db.mode, out=out)),
                ('tma', lambda out=None: engine.intersect_tma(tf, db.groups, db.mode, out=out)),
                ('crop', lambda out=None: engine.intersect_rows(tc, db.groups, db.mode, out=out, grid=grid,
                                                                pairs=pairs_list)))
        for name, fn_ in runs:
            fn = lambda t_, g_, m_, out=None, fn_=fn_: fn_(out)
            o = fn(t, db.groups, db.mode)
            for _ in range(3):
                fn(t, db.groups, db.mode, out=o)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn(t, db.groups, db.mode, out=o)
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 10
            if ref is None:
                ref = o.imat.clone()
            else:
                assert torch.equal(ref, o.imat), 'kernels disagree'
        pairs = n_img * n * n
        out.append({'frame': frame, 'instances': n, 'median_diam': diam, 'images': n_img, 'fill': fill, 'cand': cand,
                    'rows_ms': res['rows'], 'mma_ms': res['mma'], 'tma_ms': res['tma'], 'crop_join_ms': res['crop'],
                    'tma_gwordpairs_s': pairs * (frame * frame / 32.0) / res['tma'] / 1e6,
                    'rows_gpairs_s': pairs / res['rows'] / 1e6,
                    'mma_gpairs_s': pairs / res['mma'] / 1e6,
                    'mma_tops': 2.0 * pairs * frame * frame / res['mma'] / 1e9,
                    'choose_kernel': engine.choose_kernel(t, db.groups)})
        print(json.dumps(out[-1]), flush=True)


if __name__ == '__main__':
    main()
