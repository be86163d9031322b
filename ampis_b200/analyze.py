"""Instance matching and detection/segmentation scores -- drop-in for ``ampis.analyze``
(reference ampis/analyze.py:19-339) on the GPU.

The reference loops over ground-truth masks in Python and calls pycocotools' RLE.iou on
chunks of 80 predictions (analyze.py:149-164), then RLE.merge/RLE.area per match
(analyze.py:315-321).  Here one fused kernel (csrc/intersect.cu) produces, per ground-truth
mask, the first-arg-max prediction, its IoU and its intersection, so the matcher and all
det/seg scores are read off a single pass.  ``interval`` is accepted for signature
compatibility; it never affected results.
"""
from pathlib import Path

import numpy as np

from . import engine
from .structures import InstanceSet, RLEMasks, masks_to_rle, masks_to_bitmask_array  # noqa: F401


def align_instance_sets(a, b):
    """Reorder *b* to match *a* by file name, keeping common files only (analyze.py:19-51)."""
    bdict = {Path(item.filepath).name: item for item in b}
    a_ordered, b_ordered = [], []
    for item in a:
        x = bdict.get(Path(item.filepath).name, None)
        if x is not None:
            a_ordered.append(item)
            b_ordered.append(x)
    return a_ordered, b_ordered


def _check_same_size(*mask_lists):
    size = None
    for ml in mask_lists:
        for m in ml:
            s = (int(m['size'][0]), int(m['size'][1]))
            if size is None:
                size = s
            elif s != size:
                raise ValueError('masks of different image sizes cannot be compared (%s vs %s)' % (size, s))


def _rows_vs_cols(rows_rle, cols_rle, mode, dense=False):
    """One image: rows x cols through the fused kernel. Returns host arrays + device handles."""
    _check_same_size(rows_rle, cols_rle)
    table = engine.table_from_rle(list(rows_rle) + list(cols_rle))
    groups = engine.Groups.interleaved(table.device, [len(rows_rle)], [len(cols_rle)], dense=dense)
    res = engine.intersect_rows(table, groups, mode)
    return table, groups, res


def _piecewise_iou(a, b, interval=80):
    """len(a) x len(b) float64 IoU matrix (analyze.py:54-112)."""
    imax, jmax = len(a), len(b)
    if imax == 0 or jmax == 0:
        return np.zeros((imax, jmax))
    table, groups, res = _rows_vs_cols(a, b, engine.MODE_IOU, dense=True)
    return engine.iou_matrix(table, res, groups, 0).cpu().numpy()


def _match_from_rows(best_col, best_iou, n_pred, iou_thresh):
    """The bookkeeping of analyze.py:166-179 on the per-GT (arg-max, max IoU) arrays."""
    matched = best_iou > iou_thresh
    gt_idx = np.nonzero(matched)[0]
    tp = np.stack([gt_idx, best_col[matched]], axis=1).astype(int) if len(gt_idx) else np.asarray([], int)
    pred_matched = np.zeros(n_pred, bool)
    if len(gt_idx):
        pred_matched[best_col[matched]] = True
    return {'tp': tp,
            'fn': np.nonzero(~matched)[0].astype(int),
            'fp': np.nonzero(~pred_matched)[0].astype(int),
            'iou': best_iou[matched]}


def _piecewise_rle_match(gt, pred, iou_thresh=0.5, interval=80, _details=None):
    """Per-GT arg-max matching on RLE lists (analyze.py:115-181).

    For each ground-truth mask the prediction with the highest IoU is taken (first one on ties,
    only IoUs strictly above 0 count); it is a match when that IoU is strictly above
    *iou_thresh*.  Several ground-truth masks may match the same prediction."""
    G, P = len(gt), len(pred)
    if G == 0 or P == 0:
        best_col = np.full(G, -1, np.int64)
        best_iou = np.zeros(G)
        best_inter = np.zeros(G, np.int64)
        areas = None
    else:
        table, groups, res = _rows_vs_cols(gt, pred, engine.MODE_IOU)
        best_col = res.best_col[:G].cpu().numpy().astype(np.int64)
        best_iou = res.best_score[:G].cpu().numpy()
        best_inter = res.best_inter[:G].cpu().numpy().view(np.uint32).astype(np.int64)
        areas = table.areas_np()
    if _details is not None:
        _details.update(best_col=best_col, best_inter=best_inter, areas=areas)
    return _match_from_rows(best_col, best_iou, P, iou_thresh)


def rle_instance_matcher(gt, pred, iou_thresh=0.5, size=None):
    """Instance matching of two mask sets (analyze.py:184-223)."""
    gt = masks_to_rle(gt, size)
    pred = masks_to_rle(pred, size)
    return _piecewise_rle_match(gt, pred, iou_thresh)


#: legacy name of the matcher in older AMPIS releases
fast_instance_match = rle_instance_matcher


def det_seg_scores(gt, pred, iou_thresh=0.5, size=None):
    """Detection and segmentation precision/recall (analyze.py:226-339).  Raises
    ZeroDivisionError like the reference when TP+FP or TP+FN is zero."""
    gtmasks = masks_to_rle(gt, size)
    predmasks = masks_to_rle(pred, size)
    det = {}
    detection_results_ = _piecewise_rle_match(gtmasks, predmasks, iou_thresh, _details=det)
    matches_ = np.asarray(detection_results_['tp'])
    TP_det_ = len(matches_)
    FN_det_ = len(detection_results_['fn'])
    FP_det_ = len(detection_results_['fp'])
    det_precision = TP_det_ / (TP_det_ + FP_det_)
    det_recall = TP_det_ / (TP_det_ + FN_det_)
    if TP_det_:
        G = len(gtmasks)
        gi, pi = matches_[:, 0], matches_[:, 1]
        # the intersection of every matched pair was already produced by the row kernel
        seg_true_positive = det['best_inter'][gi].astype(np.int64)
        tp_gt_area = det['areas'][gi].astype(np.int64)
        tp_pred_area = det['areas'][G + pi].astype(np.int64)
    else:
        seg_true_positive = tp_gt_area = tp_pred_area = np.array([], np.int64)
    seg_false_positive = tp_pred_area - seg_true_positive
    seg_false_negative = tp_gt_area - seg_true_positive
    with np.errstate(invalid='ignore', divide='ignore'):
        seg_precision = seg_true_positive / (seg_true_positive + seg_false_positive)
        seg_recall = seg_true_positive / (seg_true_positive + seg_false_negative)
    return {'det_precision': det_precision,
            'det_recall': det_recall,
            'seg_precision': seg_precision,
            'seg_recall': seg_recall,
            'det_tp': matches_,
            'det_fn': detection_results_['fn'],
            'det_fp': detection_results_['fp'],
            'seg_tp': seg_true_positive,
            'seg_fn': seg_false_negative,
            'seg_fp': seg_false_positive,
            'det_tp_iou': detection_results_['iou']}


def merge_boxes(box1, box2):
    """Smallest [r1, r2, c1, c2] box enclosing both (analyze.py:342-376)."""
    r11, r12, c11, c12 = box1
    r21, r22, c21, c22 = box2
    return np.array([min(r11, r21), max(r12, r22), min(c11, c21), max(c12, c22)])
