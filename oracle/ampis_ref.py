"""oracle/ampis_ref.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the AMPIS Python loops on the mask-evaluation hot path
(SURVEY.md section 8a), written against ``oracle.cocomask`` the way the
reference is written against ``pycocotools.mask``.  Loop structure is kept
(per-GT loop over <=80-pred chunks, per-match merge+area, per-satellite x
per-particle merge+area) because the same functions double as the CPU
baseline timed by bench.py.  ``np.int``/``np.bool``/``np.float`` (removed from
NumPy) are spelled ``np.int64``/``np.bool_``/``np.float64``.

Inputs are plain containers (lists of RLE dicts, polygon coordinate lists,
bool arrays); the type dispatch on InstanceSet/Instances/RLEMasks is host
logic tested separately.

PARITY STATUS: the reference itself cannot be imported here (pycocotools,
detectron2, skimage, matplotlib absent; removed NumPy aliases), so parity is
pinned by analyze.py:702-728's known-answer test and by cross-formulation
checks only; everything else is "parity unpinned".

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import numpy as np

from . import cocomask as rle


# ---- ampis/analyze.py -------------------------------------------------------

def piecewise_iou(a, b, interval=80):
    """analyze.py:54-112 -- len(a) x len(b) float64 IoU matrix in <=80x80 blocks."""
    imax, jmax = len(a), len(b)
    target = np.zeros((imax, jmax))
    n_seg_a = imax // interval + int(bool(imax % interval))
    n_seg_b = jmax // interval + int(bool(jmax % interval))
    _is_crowd = np.zeros(interval, bool)
    for i in range(n_seg_a):
        i1 = interval * i
        i2 = min(i1 + interval, imax)
        a_masks = a[i1:i2]
        is_crowd = _is_crowd[:i2 - i1]
        for j in range(n_seg_b):
            j1 = interval * j
            j2 = min(j1 + interval, jmax)
            b_masks = b[j1:j2]
            target[i1:i2, j1:j2] = rle.iou(b_masks, a_masks, is_crowd).T
    return target


def piecewise_rle_match(gt, pred, iou_thresh=0.5, interval=80):
    """analyze.py:115-181 -- per-GT arg-max matching (many-to-one allowed)."""
    jmax = len(pred)
    tp, fn, iou = [], [], []
    pred_matched = np.zeros(len(pred), bool)
    n_seg_pred = jmax // interval + int(jmax % interval > 0)
    for gt_idx, gt_mask in enumerate(gt):
        iou_max = 0.
        iou_argmax = -1
        for j in range(n_seg_pred):
            j0 = interval * j
            j1 = j0 + interval
            pred_args = pred[j0:j1]
            iou_scores_ = rle.iou(pred_args, [gt_mask], [False])[:, 0]
            iou_amax_j = np.argmax(iou_scores_)
            iou_max_j = iou_scores_[iou_amax_j]
            if iou_max_j > iou_max:
                iou_max = iou_max_j
                iou_argmax = iou_amax_j + j0
        if iou_max > iou_thresh:
            tp.append([gt_idx, iou_argmax])
            iou.append(iou_max)
            pred_matched[iou_argmax] = True
        else:
            fn.append(gt_idx)
    fp = np.array([x for x, matched in enumerate(pred_matched) if not matched], int)
    return {'tp': np.asarray(tp, int),
            'fn': np.asarray(fn, int),
            'fp': np.asarray(fp, int),
            'iou': np.asarray(iou)}


def det_seg_scores(gtmasks, predmasks, iou_thresh=0.5):
    """analyze.py:226-339 on RLE lists."""
    detection_results_ = piecewise_rle_match(gtmasks, predmasks, iou_thresh)
    matches_ = np.asarray(detection_results_['tp'])
    TP_det_ = len(matches_)
    FN_det_ = len(detection_results_['fn'])
    FP_det_ = len(detection_results_['fp'])
    det_precision = TP_det_ / (TP_det_ + FP_det_)
    det_recall = TP_det_ / (TP_det_ + FN_det_)
    gtmasks_tp = [gtmasks[i[0]] for i in matches_]
    predmasks_tp = [predmasks[i[1]] for i in matches_]
    seg_true_positive = np.array([rle.area(rle.merge([m1, m2], intersect=True))
                                  for m1, m2 in zip(gtmasks_tp, predmasks_tp)], np.int64)
    tp_gt_area = np.array([rle.area(m) for m in gtmasks_tp], np.int64)
    tp_pred_area = np.array([rle.area(m) for m in predmasks_tp], np.int64)
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


# ---- ampis/applications/powder.py --------------------------------------------

def rle_satellite_match(particles, satellites, match_thresh=0.5):
    """powder.py:28-112 on RLE lists (no bbox pruning, merge+area per pair)."""
    satellite_matches, intersection_scores, satellites_unmatched = [], [], []
    particles_matched_bool = np.zeros(len(particles), dtype=np.bool_)
    for satellite_idx, satellite_mask in enumerate(satellites):
        with np.errstate(invalid='ignore', divide='ignore'):
            intersects = np.array([rle.merge_area(satellite_mask, pmask, intersect=True)
                                   for pmask in particles], np.uint32) / rle.area(satellite_mask)
        iscore_amax = np.argmax(intersects)
        iscore_max = intersects[iscore_amax]
        if iscore_max > match_thresh:
            satellite_matches.append([satellite_idx, iscore_amax])
            particles_matched_bool[iscore_amax] = True
            intersection_scores.append(iscore_max)
        else:
            satellites_unmatched.append(satellite_idx)
    particles_unmatched = np.array([i for i, matched in enumerate(particles_matched_bool)
                                    if not matched], np.int64)
    satellite_matches = np.asarray(satellite_matches, np.int64)
    satellites_unmatched = np.asarray(satellites_unmatched, np.int64)
    intersection_scores = np.asarray(intersection_scores)
    match_pairs = {x: [] for x in np.unique(satellite_matches[:, 1])}
    for match in satellite_matches:
        match_pairs[match[1]].append(match[0])
    return {'satellite_matches': satellite_matches,
            'satellites_unmatched': satellites_unmatched,
            'particles_unmatched': particles_unmatched,
            'intersection_scores': intersection_scores,
            'match_pairs': match_pairs}


def satellite_metrics(particles, n_satellites, matches):
    """powder.py:221-273 on an RLE list of particles."""
    matched_particle_idx = list(matches['match_pairs'])
    mask_areas_all = rle.area(particles)
    return {'n_satellites': n_satellites,
            'n_particles_matched': len(matched_particle_idx),
            'n_particles_all': len(particles),
            'mask_areas_matched': mask_areas_all[matched_particle_idx],
            'mask_areas_all': mask_areas_all}


def satellite_measurements(matches, n_particles_per_image, n_satellites_per_image):
    """powder.py:463-569 numerics on a list of per-image match dicts."""
    n_images = len(matches)
    n_particles_matched = sum([len(x['match_pairs'].keys()) for x in matches])
    n_particles = n_particles_matched + sum([len(x['particles_unmatched']) for x in matches])
    spp_list = []
    for m in matches:
        for v in m['match_pairs'].values():
            spp_list.append(len(v))
    spp_list = np.asarray(spp_list)
    n_satellites_matched = sum(spp_list)
    mspp = np.median(spp_list)
    n_satellites_unmatched = sum([len(x['satellites_unmatched']) for x in matches])
    sat_frac = n_particles_matched / n_particles
    unique, counts = np.unique(spp_list, return_counts=True)
    assert counts.sum() == n_particles_matched
    assert n_particles == sum(n_particles_per_image)
    assert n_satellites_matched + n_satellites_unmatched == sum(n_satellites_per_image)
    counts = counts.cumsum() / counts.sum()
    keys = ['n_images', 'n_particles', 'n_satellites', 'n_satellites_unmatched', 'n_satellited_particels',
            'sat_frac', 'mspp', 'unique_satellites_per_particle', 'counts_satellites_per_particle']
    values = [n_images, n_particles, n_satellites_matched, n_satellites_unmatched, n_particles_matched,
              sat_frac, mspp, unique, counts]
    return dict(zip(keys, values))


def psd_from_areas(areas, xvals='d_eq', yvals='cvf'):
    """powder.py:414-446 -- cumulative size distribution from a list of area arrays
    (already scaled to length^2 if a pixel size applies)."""
    if type(areas[0]) in (list, np.ndarray):
        areas = np.concatenate(areas, axis=0)
    unique, counts = np.unique(areas, return_counts=True)
    if xvals.lower() == 'd_eq':
        unique = 2 * np.sqrt(unique / np.pi)
    elif xvals.lower() != 'area':
        raise ValueError('xvals must be "d_eq" or "area"')
    if yvals.lower() == 'cvf':
        volumes = 4 / 3 * np.pi ** (-1 / 2) * unique ** (3 / 2)
        counts = volumes * counts
    elif yvals.lower() != 'counts':
        raise ValueError('yvals must be "cvf" or "counts"')
    counts = counts.cumsum()
    counts = counts / counts[-1]
    return {'x': unique, 'y': counts}


# ---- ampis/structures.py ----------------------------------------------------------

def mask_areas_rle(masks):
    """structures.py:567-571 -- RLE.area(list) -> uint32[n]."""
    return rle.area(masks)


def shoelace_area(x, y):
    """structures.py:586-610."""
    return 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))


def polygons_to_rle(polygons, size):
    """structures.py:675-677 -- first polygon of each instance, frPyObjects."""
    return [rle.frPyObjects(p, *size)[0] for p in polygons]


def rle_to_bitmask_array(masks):
    """structures.py:749-752 -- bool[n, r, c] from an RLE list."""
    return rle.decode(masks).astype(np.bool_).transpose((2, 0, 1))


def size_inliers(areas, min_thresh=100, max_thresh=100000):
    """structures.py:407-416 -- strict min < area < max."""
    inlier_min = np.ones(areas.shape, np.bool_) if min_thresh is None else areas > min_thresh
    inlier_max = np.ones(areas.shape, np.bool_) if max_thresh is None else areas < max_thresh
    return np.logical_and(inlier_min, inlier_max)


def edge_inliers(masks, size, k=1):
    """structures.py:460-468 -- masks that do not touch the k-pixel border frame."""
    r, c = size
    border = np.ones((r, c), dtype=np.bool_)
    border[k:-k, k:-k] = 0
    border = rle.encode(np.asfortranarray(border.astype(np.uint8)))
    return rle.area([rle.merge([border, x], intersect=True) for x in masks]) == 0


def rprops_basic(masks):
    """structures.py:507-508 restricted to the in-scope keys:
    area (int64), equivalent_diameter = sqrt(4*area/pi) (skimage 0.18.3
    regionprops), bbox (min_row, min_col, max_row, max_col) half-open."""
    out = {'area': [], 'equivalent_diameter': [], 'bbox': []}
    for m in masks:
        img = rle.decode(m).astype(np.int64)
        a = int(img.sum())
        out['area'].append(a)
        out['equivalent_diameter'].append(np.sqrt(4 * a / np.pi))
        rows = np.where(img.any(axis=1))[0]
        cols = np.where(img.any(axis=0))[0]
        if len(rows):
            out['bbox'].append((rows[0], cols[0], rows[-1] + 1, cols[-1] + 1))
        else:
            out['bbox'].append((0, 0, 0, 0))
    return out


# ---- ampis/data_utils.py -----------------------------------------------------------

def extract_boxes(masks, mask_mode='detectron2', box_mode='detectron2'):
    """data_utils.py:180-252."""
    if masks.ndim == 2:
        masks = masks[np.newaxis, :, :]
    else:
        if mask_mode == 'matterport':
            masks = masks.transpose((2, 0, 1))
    dtype = np.float64 if box_mode == 'detectron2' else np.int64
    boxes = np.zeros((masks.shape[0], 4), dtype=dtype)
    for i, m in enumerate(masks):
        horizontal_indicies = np.where(np.any(m, axis=0))[0]
        vertical_indicies = np.where(np.any(m, axis=1))[0]
        if horizontal_indicies.shape[0]:
            x1, x2 = horizontal_indicies[[0, -1]]
            y1, y2 = vertical_indicies[[0, -1]]
        else:
            x1, x2, y1, y2 = 0, 0, 0, 0
        if box_mode == 'detectron2':
            box = np.array([x1, y1, x2, y2], dtype=dtype)
        else:
            box = np.array([y1, y2 + 1, x1, x2 + 1], dtype=dtype)
        boxes[i] = box
    return boxes


# ---- ampis/analyze.py: boundary disagreement and pixel-class maps ----------------------------

def merge_boxes(box1, box2):
    """analyze.py:342-376 -- smallest [r1, r2, c1, c2] box enclosing both."""
    r11, r12, c11, c12 = box1
    r21, r22, c21, c22 = box2
    return np.array([min(r11, r21), max(r12, r22), min(c11, c21), max(c12, c22)])


def min_euclid(a, b):
    """analyze.py:379-413 -- for every row of a (n x 2) the smallest Euclidean distance to a row
    of b (m x 2), float64; blocked so the n x m table never materialises."""
    a = np.asarray(a, np.float64).reshape(-1, 2)
    b = np.asarray(b, np.float64).reshape(-1, 2)
    if b.shape[0] == 0:
        raise RuntimeError('min over an empty set of pixels (torch raises here too)')
    out = np.empty(a.shape[0], np.float64)
    step = max(1, (1 << 22) // max(b.shape[0], 1))
    for i in range(0, a.shape[0], step):
        d = a[i:i + step, None, :] - b[None, :, :]
        out[i:i + step] = np.sqrt((d * d).sum(axis=2)).min(axis=1)
    return out


def mask_edge_distance(gt_mask, pred_mask, gt_box, pred_box, matches):
    """analyze.py:416-499 -- per matched pair: distances from false-positive pixels to the nearest
    ground-truth pixel and from false-negative pixels to the nearest predicted pixel, inside the
    merged box [r1:r2, c1:c2]; pixel order is np.where's (row-major)."""
    fp_all, fn_all = [], []
    for gi, pi in np.asarray(matches).reshape(-1, 2):
        r1, r2, c1, c2 = merge_boxes(gt_box[gi], pred_box[pi])
        gm = rle.decode(gt_mask[gi])[r1:r2, c1:c2].astype(bool)
        pm = rle.decode(pred_mask[pi])[r1:r2, c1:c2].astype(bool)
        gt_where = np.stack(np.where(gm), axis=1)
        pred_where = np.stack(np.where(pm), axis=1)
        fp_where = np.stack(np.where(pm & ~gm), axis=1)
        fn_where = np.stack(np.where(gm & ~pm), axis=1)
        fp_all.append(min_euclid(fp_where, gt_where) if fp_where.size else np.zeros(0, np.float64))
        fn_all.append(min_euclid(fn_where, pred_where) if fn_where.size else np.zeros(0, np.float64))
    return fp_all, fn_all


def seg_perf_masks(gt_masks, pred_masks, tp_idx, mode='reduced'):
    """analyze.py:637-692 -- pixel classes of the matched pairs projected onto the frame and
    encoded: 'reduced' -> [TP, FN, FP, other], 'all' -> codes 1..7 of TP + 2 FN + 4 FP."""
    g = rle_to_bitmask_array(gt_masks)[tp_idx[:, 0]]
    p = rle_to_bitmask_array(pred_masks)[tp_idx[:, 1]]
    tp = np.logical_or.reduce(g & p, axis=0).astype(np.uint64)
    fn = np.logical_or.reduce(g & ~p, axis=0).astype(np.uint64) * 2
    fp = np.logical_or.reduce(~g & p, axis=0).astype(np.uint64) * 4
    pixel_map = tp + fn + fp
    if mode == 'all':
        masks = np.zeros((*pixel_map.shape[:2], 7), bool)
        for i in range(1, 8):
            masks[:, :, i - 1] = pixel_map == i
    else:
        masks = np.zeros((*pixel_map.shape[:2], 4), bool)
        for i, idx in enumerate([1, 2, 4]):
            masks[:, :, i] = pixel_map == idx
        masks[:, :, 3] = np.logical_or.reduce([pixel_map == i for i in [3, 5, 6, 7]], axis=0)
    return rle.encode(np.asfortranarray(masks.astype(np.uint8)))


# ---- ampis/data_utils.py: annotation images -> instances ---------------------------------------

def label_binary(ann):
    """skimage.measure.label(ann.astype(bool)) as called at data_utils.py:410: 2-D, default
    connectivity = ndim (8-connected), labels 1..n in raster order.  skimage is absent offline;
    scipy.ndimage.label with a full 3x3 structure is the same labelling with the same numbering."""
    import scipy.ndimage as ndi
    lab, _ = ndi.label(np.asarray(ann).astype(bool), structure=np.ones((3, 3), int))
    return lab


def annotations_from_label_image(ann, binary):
    """data_utils.py:409-424 -- per instance: box (extract_boxes, detectron2 mode) and RLE.encode(mask)."""
    ann = np.asarray(ann)
    if binary:
        ann = label_binary(ann)
    unique = np.unique(ann)
    if unique[0] == 0:
        unique = unique[1:]
    out = []
    for u in unique:
        mask = ann == u
        bbox = extract_boxes(mask)[0]
        out.append((bbox, rle.encode(np.asfortranarray(mask.astype(np.uint8)))))
    return out


# ---- skimage.measure.regionprops_table as called by compute_rprops (structures.py:505-508) ------
# scikit-image 0.18.3 (docker/env.yml:16) is not installable offline; the functions below restate
# its algorithms (measure/_moments.py, measure/_regionprops.py, measure/_regionprops_utils.perimeter,
# morphology/convex_hull.py, measure/_pnpoly.pxd) with numpy / scipy.  PARITY UNPINNED against a
# real skimage.

def _moments_central(image, center, order):
    calc = image.astype(float)
    for dim, dim_length in enumerate(image.shape):
        delta = np.arange(dim_length, dtype=float) - center[dim]
        powers_of_delta = delta[:, np.newaxis] ** np.arange(order + 1)
        calc = np.rollaxis(calc, dim, image.ndim)
        calc = np.dot(calc, powers_of_delta)
        calc = np.rollaxis(calc, -1, dim)
    return calc


def _perimeter4(image):
    import scipy.ndimage as ndi
    strel = np.array([[0, 1, 0], [1, 1, 1], [0, 1, 0]], np.uint8)
    image = image.astype(np.uint8)
    eroded = ndi.binary_erosion(image, strel, border_value=0)
    border = image - eroded
    weights = np.zeros(50, dtype=np.double)
    weights[[5, 7, 15, 17, 25, 27]] = 1
    weights[[21, 33]] = np.sqrt(2)
    weights[[13, 23]] = (1 + np.sqrt(2)) / 2
    pim = ndi.convolve(border, np.array([[10, 2, 10], [2, 1, 2], [10, 2, 10]]), mode='constant', cval=0)
    hist = np.bincount(pim.ravel(), minlength=50)
    return hist @ weights


def _pnpoly_grid(shape, verts):
    """grid_points_in_poly of skimage 0.18 (crossing-number test of _pnpoly.pxd, x = row, y = col)."""
    xp, yp = verts[:, 0], verts[:, 1]
    m, n = np.mgrid[0:shape[0], 0:shape[1]]
    x, y = m.astype(float), n.astype(float)
    c = np.zeros(shape, bool)
    j = len(xp) - 1
    for i in range(len(xp)):
        cond = ((yp[i] <= y) & (y < yp[j])) | ((yp[j] <= y) & (y < yp[i]))
        with np.errstate(divide='ignore', invalid='ignore'):
            xi = (xp[j] - xp[i]) * (y - yp[i]) / (yp[j] - yp[i]) + xp[i]
        c ^= cond & (x < xi)
        j = i
    return c


def _convex_hull_image(image):
    from scipy.spatial import ConvexHull
    if not image.any():
        return np.zeros(image.shape, bool)
    rows, cols = np.nonzero(image)
    coords = []
    for r in np.unique(rows):                    # possible_hull: first / last pixel of every row and column
        cc = cols[rows == r]
        coords += [(r, cc.min()), (r, cc.max())]
    for c in np.unique(cols):
        rr = rows[cols == c]
        coords += [(rr.min(), c), (rr.max(), c)]
    coords = np.array(coords, float)
    offsets = np.array([[-0.5, 0], [0.5, 0], [0, -0.5], [0, 0.5]])
    coords = (coords[:, None, :] + offsets).reshape(-1, 2)
    coords = np.unique(coords, axis=0)
    hull = ConvexHull(coords)
    verts = hull.points[hull.vertices]
    return _pnpoly_grid(image.shape, verts)


def regionprops_one(mask):
    """Properties of the single region `mask.astype(int)` defines (label 1), as regionprops_table
    would report them; {} for an empty mask."""
    mask = np.asarray(mask, bool)
    if not mask.any():
        return {}
    rows, cols = np.nonzero(mask)
    r0, r1, c0, c1 = rows.min(), rows.max() + 1, cols.min(), cols.max() + 1
    img = mask[r0:r1, c0:c1]
    area = int(img.sum())
    M = _moments_central(img.astype(np.uint8), (0, 0), 3)
    local_centroid = (M[1, 0] / M[0, 0], M[0, 1] / M[0, 0])
    mu = _moments_central(img.astype(np.uint8), local_centroid, 3)
    mu0 = mu[0, 0]
    T = np.array([[mu[0, 2] / mu0, -mu[1, 1] / mu0], [-mu[1, 1] / mu0, mu[2, 0] / mu0]])
    ev = np.linalg.eigvalsh(T)
    ev = np.clip(ev, 0, None, out=ev)
    l1, l2 = sorted(ev, reverse=True)
    a, b, b, c = T.flat
    if a - c == 0:
        orientation = -np.pi / 4. if b < 0 else np.pi / 4.
    else:
        orientation = 0.5 * np.arctan2(-2 * b, c - a)
    convex_area = int(_convex_hull_image(img).sum())
    return {'area': area, 'bbox': (int(r0), int(c0), int(r1), int(c1)), 'bbox_area': int(img.size),
            'centroid': (float(rows.mean()), float(cols.mean())),
            'local_centroid': (float(local_centroid[0]), float(local_centroid[1])),
            'convex_area': convex_area, 'eccentricity': 0. if l1 == 0 else float(np.sqrt(1 - l2 / l1)),
            'equivalent_diameter': float(np.sqrt(4 * area / np.pi)), 'extent': area / img.size,
            'major_axis_length': float(4 * np.sqrt(l1)), 'minor_axis_length': float(4 * np.sqrt(l2)),
            'orientation': float(orientation), 'perimeter': float(_perimeter4(img)),
            'solidity': area / convex_area, 'label': 1}


def poly2mask(polygons, size):
    """structures.py:693-715 -- skimage.draw.polygon2mask of every [x0,y0,x1,y1,...] polygon:
    draw.polygon's integer bounding range, skimage 0.18 _pnpoly crossing rule on (x = col, y = row)."""
    out = np.zeros((len(polygons), size[0], size[1]), bool)
    for k, p in enumerate(polygons):
        p = np.asarray(p, float)
        c, r = p[0::2], p[1::2]                       # vertex columns (x) and rows (y)
        minr, maxr = int(max(0, r.min())), min(size[0] - 1, int(np.ceil(r.max())))
        minc, maxc = int(max(0, c.min())), min(size[1] - 1, int(np.ceil(c.max())))
        if maxr < minr or maxc < minc:
            continue
        yy, xx = np.mgrid[minr:maxr + 1, minc:maxc + 1]
        x, y = xx.astype(float), yy.astype(float)
        inside = np.zeros(x.shape, bool)
        j = len(c) - 1
        for i in range(len(c)):
            cond = ((r[i] <= y) & (y < r[j])) | ((r[j] <= y) & (y < r[i]))
            with np.errstate(divide='ignore', invalid='ignore'):
                xi = (c[j] - c[i]) * (y - r[i]) / (r[j] - r[i]) + c[i]
            inside ^= cond & (x < xi)
            j = i
        out[k, minr:maxr + 1, minc:maxc + 1] = inside
    return out
