"""Generates tests/golden/*.npz from the reference's shipped data fixtures and
the CPU oracle.  Run in the build container only (it reads /root/reference,
which does not exist on the GPU box):

    python tests/golden/make_golden.py

Inputs (all under /root/reference/examples):
  powder/data/{sample_particle_outputs,particle-results,satellite-results}.pickle
      -- Mask R-CNN predictions, masks as COCO-compressed RLE produced by the
         REAL pycocotools.encode (data_utils.py:275): golden vectors for the
         string codec / encode canonical form.
  powder/data/via_2.0.8/via_powder_particle_masks_validation.json -- GT polygons
  spheroidite/data/sample-spheroidite-results.pickle

Expected outputs are produced by oracle/ (no real pycocotools exists here), so
they freeze the oracle's behaviour ("parity unpinned" beyond the codec and the
analyze.py:702-728 known-answer test).
"""
import json
import os
import pickle
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ampis_ref as R          # noqa: E402
from oracle import cocomask as rle         # noqa: E402

REF = '/root/reference/examples'
OUT = os.path.dirname(os.path.abspath(__file__))


class _Inst:
    def __setstate__(self, s):
        self.__dict__.update(s)


class _U(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith('detectron2') and name == 'Instances':
            return _Inst
        if module.split('.')[0] == 'numpy' or module in ('builtins', 'collections', 'copyreg', '_codecs'):
            return super().find_class(module, name)
        raise pickle.UnpicklingError('%s.%s' % (module, name))


def load_preds(path):
    out = {}
    for e in _U(open(path, 'rb')).load():
        inst = e['pred']['instances']
        out[os.path.basename(str(e['file_name']))] = {
            'size': tuple(inst._image_size), 'masks': inst._fields['pred_masks'],
            'boxes': np.asarray(inst._fields['pred_boxes']), 'scores': np.asarray(inst._fields['scores'])}
    return out


def pack_strings(masks):
    """list of RLE dicts -> (uint8 blob, int64 offsets[n+1])"""
    bs = [m['counts'] for m in masks]
    off = np.zeros(len(bs) + 1, np.int64)
    off[1:] = np.cumsum([len(b) for b in bs])
    return np.frombuffer(b''.join(bs), np.uint8).copy(), off


def main():
    pa = load_preds(REF + '/powder/data/sample_particle_outputs.pickle')
    pb = load_preds(REF + '/powder/data/particle-results.pickle')
    sb = load_preds(REF + '/powder/data/satellite-results.pickle')
    sp = load_preds(REF + '/spheroidite/data/sample-spheroidite-results.pickle')

    # ---- 1. instance matching: sample_particle_outputs ("gt") vs particle-results ("pred")
    g = {}
    names = sorted(set(pa) & set(pb))
    g['names'] = np.array(names)
    for k, nm in enumerate(names):
        gt, pr = pa[nm]['masks'], pb[nm]['masks']
        g['%d_size' % k] = np.array(pa[nm]['size'])
        g['%d_gt_blob' % k], g['%d_gt_off' % k] = pack_strings(gt)
        g['%d_pr_blob' % k], g['%d_pr_off' % k] = pack_strings(pr)
        res = R.det_seg_scores(gt, pr, 0.5)
        for key in ('det_tp', 'det_fn', 'det_fp', 'seg_tp', 'seg_fn', 'seg_fp', 'det_tp_iou',
                    'seg_precision', 'seg_recall'):
            g['%d_%s' % (k, key)] = np.asarray(res[key])
        g['%d_det_pr' % k] = np.array([res['det_precision'], res['det_recall']])
        iou = R.piecewise_iou(gt, pr)
        nz = np.argwhere(iou > 0)
        g['%d_iou_nz_idx' % k] = nz.astype(np.int32)
        g['%d_iou_nz_val' % k] = iou[nz[:, 0], nz[:, 1]]
        g['%d_gt_area' % k] = rle.area(gt)
        g['%d_pr_area' % k] = rle.area(pr)
        # counts at the 10 COCO thresholds
        ths = np.arange(0.5, 1.0, 0.05)
        cnt = []
        for t in ths:
            m = R.piecewise_rle_match(gt, pr, t)
            cnt.append([len(m['tp']), len(m['fp']), len(m['fn'])])
        g['%d_thr_counts' % k] = np.array(cnt, np.int64)
        g['thresholds'] = ths
        print('match', nm, len(gt), len(pr), cnt[0])
    np.savez_compressed(OUT + '/powder_match.npz', **g)

    # ---- 2. satellites: satellite-results vs particle-results
    s = {}
    names = sorted(set(sb) & set(pb))
    s['names'] = np.array(names)
    for k, nm in enumerate(names):
        part, sat = pb[nm]['masks'], sb[nm]['masks']
        s['%d_size' % k] = np.array(pb[nm]['size'])
        s['%d_sat_blob' % k], s['%d_sat_off' % k] = pack_strings(sat)
        s['%d_part_ref' % k] = np.array(nm)   # particle strings live in powder_match.npz (same image)
        res = R.rle_satellite_match(part, sat, 0.5)
        for key in ('satellite_matches', 'satellites_unmatched', 'particles_unmatched', 'intersection_scores'):
            s['%d_%s' % (k, key)] = np.asarray(res[key])
        print('sat', nm, len(part), len(sat), len(res['satellite_matches']))
    np.savez_compressed(OUT + '/powder_satellite.npz', **s)

    # ---- 3. polygons (GT) -> RLE strings
    j = json.load(open(REF + '/powder/data/via_2.0.8/via_powder_particle_masks_validation.json'))
    p = {}
    for k, annos in enumerate(j['_via_img_metadata'].values()):
        size = annos['file_attributes'].get('Size (width, height)', None)
        width, height = (int(x) for x in size.split(', '))
        polys = []
        for obj in annos['regions']:
            sh = obj['shape_attributes']
            poly = [c for x, y in zip(sh['all_points_x'], sh['all_points_y']) for c in (x + 0.5, y + 0.5)]
            polys.append(np.asarray(poly, np.float64))
        rl = R.polygons_to_rle([[q] for q in polys], (height, width))
        p['%d_size' % k] = np.array([height, width])
        p['%d_poly_xy' % k] = np.concatenate(polys)
        off = np.zeros(len(polys) + 1, np.int64)
        off[1:] = np.cumsum([len(q) for q in polys])
        p['%d_poly_off' % k] = off
        p['%d_rle_blob' % k], p['%d_rle_off' % k] = pack_strings(rl)
        p['%d_area' % k] = rle.area(rl)
        p['%d_name' % k] = np.array(annos['filename'])
        print('poly', annos['filename'], len(polys))
    p['n_images'] = np.array(k + 1)
    np.savez_compressed(OUT + '/powder_polygons.npz', **p)

    # ---- 4. spheroidite: measurements on 3 images
    m = {}
    names = sorted(sp)[:3]
    m['names'] = np.array(names)
    for k, nm in enumerate(names):
        masks = sp[nm]['masks']
        m['%d_size' % k] = np.array(sp[nm]['size'])
        m['%d_blob' % k], m['%d_off' % k] = pack_strings(masks)
        m['%d_area' % k] = rle.area(masks)
        bm = R.rle_to_bitmask_array(masks)
        m['%d_boxes_d2' % k] = R.extract_boxes(bm)
        m['%d_boxes_mp' % k] = R.extract_boxes(bm, box_mode='matterport')
        rp = R.rprops_basic(masks)
        m['%d_deq' % k] = np.asarray(rp['equivalent_diameter'])
        m['%d_edge_inliers' % k] = R.edge_inliers(masks, sp[nm]['size'], 1)
        m['%d_size_inliers' % k] = R.size_inliers(rle.area(masks), 100, 100000)
        m['%d_rlebbox' % k] = np.stack([rle.to_bbox(x) for x in masks])
        print('sph', nm, len(masks))
    np.savez_compressed(OUT + '/spheroidite_measure.npz', **m)


if __name__ == '__main__':
    main()
