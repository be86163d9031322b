"""Mask containers and converters -- drop-in for ``ampis.structures`` (reference
ampis/structures.py) with the mask arithmetic running on the GPU.

Same names, argument order, defaults, return dtypes and exceptions as the reference; the
dispatch on ``type(x) ==`` (not isinstance) is deliberate, it is what the reference does
(structures.py:57-81, 556-583, 663-690, 736-774).
"""
import colorsys
import copy
from pathlib import Path
from typing import List, Union

import numpy as np
import pandas as pd
import torch

from . import engine
from .containers import BitMasks, Boxes, Instances, PolygonMasks


def random_colors(n, seed, bright=True):
    """n distinguishable RGB colours in a seeded random order (reference ampis/visualize.py:19-56; the
    InstanceSet readers attach them as the ``colors`` field): evenly spaced hues, full saturation,
    value 1.0 (0.7 when not *bright*), shuffled by RandomState(seed)."""
    value = 1.0 if bright else 0.7
    palette = [colorsys.hsv_to_rgb(k / n, 1, value) for k in range(n)]
    np.random.RandomState(seed=seed).shuffle(palette)
    return np.asarray(palette)


def _is_bool_selector(item):
    """Boolean masks in the forms RLEMasks.__getitem__ accepts: torch bool tensor, numpy bool array,
    list whose first element is a bool."""
    if type(item) == torch.BoolTensor or (type(item) == torch.Tensor and item.dtype == torch.bool):
        return True
    if type(item) == np.ndarray:
        return item.dtype == np.bool_
    return type(item) == list and type(item[0]) == bool


class RLEMasks(object):
    """List of COCO RLE dicts that indexes like detectron2's mask containers (reference
    structures.py:24-95): int, slice, boolean mask (same length) or a sequence of indices."""

    def __init__(self, rle):
        super().__init__()
        self.rle = rle

    def __getitem__(self, item: Union[int, slice, List[int], List[bool], torch.BoolTensor, np.ndarray]):
        if type(item) in (int, slice):
            return RLEMasks(self.rle[item])          # an int wraps ONE dict, like the reference (quirk B.13)
        if _is_bool_selector(item):
            if type(item) in (np.ndarray, list):
                assert len(item) == len(self)
            return RLEMasks([m for m, keep in zip(self.rle, item) if keep])
        return RLEMasks([self.rle[k] for k in item])

    def __len__(self):
        return len(self.rle)


def _parse_hfw(value):
    """'103.6 um' -> (103.6, 'um'); a bare number -> (number, None); None -> (None, None); anything else
    is kept as it is (reference structures.py:294-305)."""
    if value is None:
        return None, None
    try:
        return float(value), None
    except ValueError:
        parts = value.split(' ')
        if len(parts) == 2:
            return float(parts[0]), parts[1]
        return value, None


class InstanceSet(object):
    """Everything AMPIS keeps about the instances of one image (reference structures.py:98-533): the
    ``Instances`` (masks, boxes, class_idx, scores, colors), where they came from (``filepath``,
    ``dataset_class``, ``pred_or_gt``, ``mask_format``), the physical scale (``HFW``, ``HFW_units``)
    and the measured ``rprops``."""

    def __init__(self, mask_format=None, bbox_mode=None, filepath=None, annotations=None, instances=None, img=None,
                 dataset_class=None, pred_or_gt=None, HFW=None, HFW_units=None, randomstate=None):
        super().__init__()
        self.mask_format, self.bbox_mode = mask_format, bbox_mode
        self.filepath, self.img = filepath, img
        self.annotations, self.instances = annotations, instances
        self.dataset_class, self.pred_or_gt = dataset_class, pred_or_gt
        self.HFW, self.HFW_units = HFW, HFW_units
        self.rprops = None
        self.colors = None
        self.randomstate = np.random.randint(2 ** 32 - 1) if randomstate is None else randomstate

    def _attach(self, instances):
        self.instances = instances
        self.instances.colors = random_colors(len(instances), self.randomstate)

    def read_from_ddict(self, ddict, inplace=True):
        """Fill the set from a ground-truth data dict (reference structures.py:203-309): RLE dicts
        become RLEMasks, bool arrays BitMasks, coordinate lists PolygonMasks; ``HFW`` strings like
        '103.6 um' are split into value and unit."""
        annos = ddict['annotations']
        self.pred_or_gt, self.filepath, self.mask_format = 'gt', Path(ddict['file_name']), ddict['mask_format']
        segs = [a['segmentation'] for a in annos]
        kind = type(segs[0])
        if kind == dict:
            masks = RLEMasks(segs)
        elif kind == np.ndarray:
            if segs[0].dtype == np.bool_:
                masks = BitMasks(np.stack(segs))
        else:
            masks = PolygonMasks(segs)
        self._attach(Instances((ddict['height'], ddict['width']), masks=masks,
                               boxes=np.stack([a['bbox'] for a in annos]),
                               class_idx=np.asarray([a['category_id'] for a in annos], np.int64)))
        self.dataset_class = ddict.get('dataset_class', None)
        self.HFW, self.HFW_units = _parse_hfw(ddict.get('HFW', None))
        return None if inplace else self

    def read_from_model_out(self, outs, inplace=True):
        """Fill the set from ``data_utils.format_outputs`` output (reference structures.py:312-371).
        A dataset name like 'powder_Validation' yields dataset_class 'Validation'."""
        self.pred_or_gt, self.mask_format, self.filepath = 'pred', 'bitmask', outs['file_name']
        self.dataset_class = outs['dataset'].split('_')[-1]
        pred = outs['pred']['instances']
        self._attach(Instances(pred.image_size, masks=RLEMasks(pred.pred_masks), boxes=pred.pred_boxes,
                               class_idx=pred.pred_classes, scores=pred.scores))
        return None if inplace else self

    def filter_mask_size(self, min_thresh=100, max_thresh=100000, to_rle=False):
        """New ``Instances`` holding the instances with min_thresh < area < max_thresh (both strict,
        ``None`` disables a side); the set itself is not modified (reference structures.py:374-442)."""
        masks = self.instances.masks
        if to_rle:
            masks = RLEMasks(masks_to_rle(masks, self.instances.image_size))
        areas = mask_areas(masks)
        keep = np.ones(areas.shape, np.bool_)
        if min_thresh is not None:
            keep &= areas > min_thresh
        if max_thresh is not None:
            keep &= areas < max_thresh
        if type(masks) == PolygonMasks:
            masks = PolygonMasks([p for p, k in zip(masks.polygons, keep) if k])
        else:
            masks = masks[keep]
        fields = {name: (masks if name == 'masks' else value[keep]) for name, value in self.instances._fields.items()}
        return Instances(self.instances.image_size, **fields)

    def remove_edge_instances(self, k=1):
        """Drop instances that touch the k-pixel image border, in place (reference
        structures.py:445-469).  The reference intersects every mask with an RLE border frame
        (``border[k:-k, k:-k] = 0``); a mask meets that frame iff its tight bounding box does,
        so the GPU measurement pass answers it without any merge.  ``k=0`` keeps the reference
        quirk: the slice is empty, the frame is the whole image and every non-empty mask goes."""
        r, c = self.instances.image_size
        rle = masks_to_rle(self.instances.masks, (r, c))
        area, bb = engine.measure_rle(rle)
        # ``border[k:-k, k:-k] = 0`` zeroes nothing when the slice is empty
        if k <= 0 or r - 2 * k <= 0 or c - 2 * k <= 0:
            touches = area > 0
        else:
            # frame = complement of the zeroed interior [k, r-k) x [k, c-k)
            touches = (area > 0) & ((bb[:, 0] < k) | (bb[:, 2] >= c - k) | (bb[:, 1] < k) | (bb[:, 3] >= r - k))
        inlier_instances = ~touches
        self.instances = self.instances[inlier_instances]

    #: region properties compute_rprops derives from the GPU measurements (skimage names)
    RPROPS_GPU_KEYS = ('area', 'bbox', 'bbox_area', 'centroid', 'local_centroid', 'convex_area', 'eccentricity',
                       'equivalent_diameter', 'extent', 'label', 'major_axis_length', 'minor_axis_length',
                       'orientation', 'perimeter', 'solidity')

    def compute_rprops(self, keys=None, return_df=False):
        """Region properties per mask (reference structures.py:474-514, which decodes every mask to a
        full int64 frame and runs skimage.measure.regionprops_table on it).  Here the GPU delivers
        exact integer measurements from the run table and the packed bounding-box windows
        (csrc/rprops.cu: raw moments, perimeter histogram, convex-hull pixel count) and the
        properties are formed on the host with skimage 0.18.3's formulas.  Default keys are the
        reference's; cells are 1-element arrays (empty for an empty mask) as regionprops_table
        returns them; tuple-valued properties become ``key-0``, ``key-1`` ... columns."""
        if keys is None:
            keys = ['area', 'equivalent_diameter', 'major_axis_length', 'perimeter', 'solidity', 'orientation']
        unsupported = [k for k in keys if k not in self.RPROPS_GPU_KEYS]
        if unsupported:
            raise NotImplementedError('region properties %s are not part of the GPU path (supported: %s)'
                                      % (unsupported, list(self.RPROPS_GPU_KEYS)))
        rle = masks_to_rle(self.instances.masks, self.instances.image_size)
        if type(rle) == dict:
            rle = [rle]
        props = region_properties(rle)
        rows = []
        for p in props:
            row = {}
            for k in keys:
                v = p.get(k)
                if isinstance(v, tuple):
                    for j, x in enumerate(v):
                        row['%s-%d' % (k, j)] = np.array([x])
                elif v is None:            # empty mask: regionprops finds no region
                    width = {'bbox': 4, 'centroid': 2, 'local_centroid': 2}.get(k, 0)
                    empty = np.array([], np.int64 if k in ('area', 'bbox', 'bbox_area', 'convex_area', 'label')
                                     else np.float64)
                    if width:
                        for j in range(width):
                            row['%s-%d' % (k, j)] = empty
                    else:
                        row[k] = empty
                else:
                    row[k] = np.array([v])
            rows.append(row)
        df = pd.DataFrame(rows)
        df['class_idx'] = self.instances.class_idx
        self.rprops = df
        if return_df:
            return self.rprops

    def copy(self):
        return copy.deepcopy(self)


def mask_areas(masks):
    """Area in pixels of each mask (reference structures.py:536-583): ndarray -> uint64 sums,
    PolygonMasks -> shoelace float64, RLE -> uint32 pixel counts (GPU), containers recurse."""
    kind = type(masks)
    if kind == np.ndarray:
        if masks.dtype == np.bool_ and masks.ndim == 3 and masks.size:
            return engine.bool_area_bbox(masks)[0].astype(np.uint)
        return masks.sum(axis=(1, 2), dtype=np.uint)
    if kind == PolygonMasks:
        return np.asarray([_shoelace_area(ring[0][::2], ring[0][1::2]) for ring in masks.polygons])
    if kind == RLEMasks or (kind == list and type(masks[0]) == dict):
        return engine.measure_rle(masks.rle if kind == RLEMasks else masks)[0]
    if kind in (Instances, InstanceSet):
        return mask_areas(masks.masks if kind == Instances else masks.instances)
    if kind == list:
        return [mask_areas(m) for m in masks]
    raise NotImplementedError('Not implemented for type {}'.format(kind))


def _shoelace_area(x, y):
    """Polygon area from vertex coordinates (reference structures.py:586-610)."""
    return 0.5 * np.abs(np.dot(x, np.roll(y, 1)) - np.dot(y, np.roll(x, 1)))


def boxes_to_array(boxes):
    """n x 4 ndarray from an ndarray, a list of 4-sequences or detectron2 ``Boxes`` (reference
    structures.py:613-639); other types give None, as in the reference."""
    if type(boxes) == np.ndarray:
        return boxes
    if type(boxes) == list:
        assert len(boxes[0]) == 4
        return np.asarray(boxes)
    if type(boxes) == Boxes:
        return boxes.tensor.to('cpu').numpy()


def masks_to_rle(masks, size=None):
    """Any supported mask container -> list of COCO RLE dicts (reference structures.py:643-690).
    RLE input is returned as it is.  Polygons need *size* and are rasterised by the GPU restatement of
    pycocotools' rleFrPoly -- the first polygon of each instance only, as ``RLE.frPyObjects(p, *size)[0]``
    does -- and compressed by the GPU string encoder."""
    kind = type(masks)
    if kind == list:
        if type(masks[0]) == dict:
            return masks
        if type(masks[0]) == list:
            raise NotImplementedError('):')
    if kind == RLEMasks:
        return masks.rle
    if kind == PolygonMasks:
        assert size is not None
        h, w = int(size[0]), int(size[1])
        rings = [np.asarray(p[0], np.float64) for p in masks.polygons]
        cnt, cnt_off, cnt_len, _, _ = engine.polygons_to_counts(rings, h, w)
        return [{'size': [h, w], 'counts': c} for c in engine.counts_to_strings(cnt, cnt_off, cnt_len, len(rings))]
    if kind in (Instances, InstanceSet):
        inst = masks if kind == Instances else masks.instances
        return masks_to_rle(inst.masks, inst.image_size)
    raise NotImplementedError('cannot convert mask type {} to RLE'.format(masks))


def _rle_to_bool(rle):
    t = engine.table_from_rle(rle)
    h, w = rle[0]['size']
    for m in rle:
        if list(m['size']) != [h, w]:
            raise ValueError('all masks must share one image size')
    return engine.to_host(engine.unpack_bool(t, np.arange(len(rle)), int(h), int(w)))


def _poly2mask(masks, size):
    """list of [x0,y0,x1,y1,...] polygons -> bool[n, r, c] with skimage.draw.polygon2mask's rule
    (reference structures.py:693-715), rasterised on the GPU (csrc/poly.cu polygon2mask_kernel)."""
    return engine.polygons_to_bool([np.asarray(p, np.float64).ravel() for p in masks], size[0], size[1])


def masks_to_bitmask_array(masks, size=None):
    """Anything -> bool[n_mask, r, c] (reference structures.py:717-774).  RLE input is decoded
    and transposed on the GPU.  Polygon input goes through ``_poly2mask`` like the reference
    (first polygon of every instance, skimage's point-in-polygon rule -- NOT the pycocotools
    rasteriser ``masks_to_rle`` uses; the two differ on boundary pixels, as in the reference)."""
    dtype = type(masks)
    if dtype == np.ndarray:
        assert masks.dtype == np.bool_
        return masks
    elif dtype == PolygonMasks:
        assert size is not None
        return _poly2mask([p[0] for p in masks.polygons], size)
    elif dtype == list:
        if type(masks[0]) == dict:
            return _rle_to_bool(masks)
        elif type(masks[0]) == list or type(masks[0]) == np.ndarray:
            assert size is not None
            return _poly2mask(masks, size)
        else:
            raise NotImplementedError
    elif dtype == RLEMasks:
        if type(masks.rle) == dict:      # RLEMasks[int] wraps one dict (quirk B.13)
            return _rle_to_bool([masks.rle])
        return _rle_to_bool(masks.rle)
    elif dtype == InstanceSet:
        return masks_to_bitmask_array(masks.instances.masks, masks.instances.image_size)
    elif dtype == Instances:
        return masks_to_bitmask_array(masks.masks, masks.image_size)
    else:
        raise NotImplementedError


_PERIMETER_CODES = [5, 7, 15, 17, 25, 27, 21, 33, 13, 23]


def region_properties(rle):
    """skimage 0.18.3 region properties of every mask of an RLE list (one region per mask, as
    ``regionprops_table(mask.astype(int))`` sees it) from the exact GPU measurements.  Returns one
    dict per mask ({} for an empty mask)."""
    m = engine.region_measurements(rle)
    weights = np.zeros(50, dtype=np.double)
    weights[[5, 7, 15, 17, 25, 27]] = 1
    weights[[21, 33]] = np.sqrt(2)
    weights[[13, 23]] = (1 + np.sqrt(2)) / 2
    out = []
    for i in range(len(rle)):
        n = int(m['area'][i])
        if n == 0:
            out.append({})
            continue
        x0, y0, x1, y1 = (int(v) for v in m['bbox'][i])
        N, Sx, Sy, Sxx, Syy, Sxy = (int(v) for v in m['moments'][i])
        # central second moments times N, exact integers: N*mu20 = N*Syy - Sy^2, ...
        n_rr, n_cc, n_rc = N * Syy - Sy * Sy, N * Sxx - Sx * Sx, N * Sxy - Sx * Sy
        a, c, b = n_cc / (N * N), n_rr / (N * N), -(n_rc / (N * N))    # inertia tensor [[a, b], [b, c]]; -0.0 kept
        half_tr, half_diff = (a + c) / 2, (a - c) / 2
        l1 = half_tr + np.sqrt(half_diff * half_diff + b * b)
        det = (n_cc * n_rr - n_rc * n_rc) / (N ** 4)                    # exact numerator: no cancellation
        l2 = max(det / l1, 0.0) if l1 > 0 else 0.0
        if n_cc == n_rr:
            orientation = -np.pi / 4. if b < 0 else np.pi / 4.
        else:
            orientation = 0.5 * np.arctan2(-2 * b, c - a)
        hist = np.zeros(50, np.int64)
        hist[_PERIMETER_CODES] = m['perimeter_hist'][i]
        convex_area = int(m['convex_area'][i])
        bbox_area = (y1 - y0 + 1) * (x1 - x0 + 1)
        out.append({'area': n, 'bbox': (y0, x0, y1 + 1, x1 + 1), 'bbox_area': bbox_area,
                    'centroid': (Sy / N, Sx / N), 'local_centroid': ((Sy - N * y0) / N, (Sx - N * x0) / N),
                    'convex_area': convex_area, 'eccentricity': 0. if l1 == 0 else float(np.sqrt(1 - l2 / l1)),
                    'equivalent_diameter': float(np.sqrt(4 * n / np.pi)), 'extent': n / bbox_area,
                    'major_axis_length': float(4 * np.sqrt(l1)), 'minor_axis_length': float(4 * np.sqrt(l2)),
                    'orientation': float(orientation), 'perimeter': float(hist @ weights),
                    'solidity': n / convex_area, 'label': 1})
    return out
