"""Drop-in for the parts of ``ampis.data_utils`` that sit on, or directly either side of, the
mask-evaluation path (reference ampis/data_utils.py): ``extract_boxes`` (:180-252), the prediction
compressor ``compress_pred`` / ``format_outputs`` (:255-310) and the ground-truth loader
``get_ddicts`` (:313-530).  The trainer and its evaluation hook (:37-177) are detectron2 training
glue and outside this package's scope."""
import json
from pathlib import Path

import numpy as np
import torch

from . import engine
from .containers import BoxMode


def extract_boxes(masks, mask_mode='detectron2', box_mode='detectron2'):
    """Tight bounding boxes of boolean masks.

    masks: bool array, ``n x r x c`` (mask_mode 'detectron2'), ``r x c x n`` ('matterport') or a
    single ``r x c`` mask.  Returns ``n x 4``: float ``[x1, y1, x2, y2]`` with inclusive maxima
    (box_mode 'detectron2') or int ``[y1, y2+1, x1, x2+1]`` ('matterport'); empty masks give
    zeros.  The per-mask min/max reduction runs on the GPU (csrc/rle_paint.cu)."""
    if masks.ndim == 2:
        masks = masks[np.newaxis, :, :]
    else:
        if mask_mode == 'matterport':
            masks = masks.transpose((2, 0, 1))
    dtype = np.float64 if box_mode == 'detectron2' else np.int64
    boxes = np.zeros((masks.shape[0], 4), dtype=dtype)
    if masks.shape[0] == 0:
        return boxes
    # bool input is uploaded as it is; anything else is reduced to "non-zero" first (numpy truthiness)
    area, bb = engine.bool_area_bbox(masks if masks.dtype in (np.bool_, np.uint8) else masks != 0)
    ne = area > 0
    x1, y1, x2, y2 = bb[ne, 0], bb[ne, 1], bb[ne, 2], bb[ne, 3]
    if box_mode == 'detectron2':
        boxes[ne] = np.stack([x1, y1, x2, y2], axis=1)
    else:
        boxes[ne] = np.stack([y1, y2 + 1, x1, x2 + 1], axis=1)
    return boxes


def _to_numpy(x):
    if hasattr(x, 'tensor'):          # detectron2 Boxes
        x = x.tensor
    if hasattr(x, 'detach'):
        x = x.detach().to('cpu').numpy()
    return np.asarray(x)


def compress_pred(pred):
    """Predicted bitmasks -> COCO RLE, every other field -> numpy, in place (data_utils.py:255-279).
    The reference encodes mask by mask with pycocotools on the CPU; here the whole ``n x h x w``
    stack is packed, run-length encoded and string-encoded on the GPU in one go."""
    masks = pred.pred_masks
    if hasattr(masks, 'detach'):                       # torch bool [n, h, w]: encoded on the device it lives on
        masks = masks.detach()
        pred.pred_masks = engine.encode_bool(masks.to(torch.bool)) if len(masks) else []
    else:
        masks = np.stack([_to_numpy(x) for x in masks]) if len(masks) else np.zeros((0, 1, 1), bool)
        pred.pred_masks = engine.encode_bool(masks.astype(np.bool_)) if len(masks) else []
    pred.pred_boxes = _to_numpy(pred.pred_boxes)
    pred.scores = _to_numpy(pred.scores)
    pred.pred_classes = _to_numpy(pred.pred_classes)
    return pred


def format_outputs(filename, dataset, pred):
    """{'file_name', 'dataset', 'pred'} with the predictions compressed in place (data_utils.py:282-310)."""
    compress_pred(pred['instances'])
    return {'file_name': filename, 'dataset': dataset, 'pred': pred}


def _imread(path, as_gray=False):
    """skimage.io.imread stand-in (skimage is not a dependency here): PIL -> numpy."""
    from PIL import Image
    im = Image.open(str(path))
    if as_gray:
        im = im.convert('L')
    return np.asarray(im)


def _record(idx, image_path, annotation_file, height, width, mask_format, dataset_class, annotations, **extra):
    """One detectron2-style data dict (the key set of data_utils.py:398-404 / 452-459 / 503-509)."""
    d = {'file_name': str(image_path), 'annotation_file': annotation_file, 'height': height, 'width': width,
         'mask_format': mask_format, 'image_id': idx}
    d.update(extra)
    d['dataset_class'] = dataset_class
    d['annotations'] = annotations
    d['num_instances'] = len(annotations)
    return d


def _instance(bbox, segmentation):
    return {'bbox': bbox, 'bbox_mode': BoxMode.XYXY_ABS, 'segmentation': segmentation, 'category_id': 0}


def _from_annotation_images(binary, im_root, ann_root, pattern, dataset_class, cwd):
    """'binary' / 'label': one annotation image (or .npy) per picture, matched by file stem."""
    out = []
    for idx, picture in enumerate(Path(im_root).glob(pattern)):
        hits = list(Path(ann_root).glob('*{}*'.format(picture.stem)))
        n = len(hits)
        assert n == 1, f'There must be exactly 1 annotation file for, {picture.name}, but {n} were found'
        ann_path = hits[0].relative_to(cwd)
        ann = np.load(str(ann_path)) if ann_path.suffix == '.npy' else _imread(ann_path)
        # labelling, boxes and RLE of every instance from ONE pass over the image on the GPU (csrc/label.cu)
        rles, boxes = engine.label_image_to_instances(ann, binary=binary)
        instances = [_instance(b.astype(np.float64), m) for m, b in zip(rles, boxes)]     # box = x1, y1, x2, y2
        out.append(_record(idx, picture.relative_to(cwd), str(ann_path), ann.shape[0], ann.shape[1], 'bitmask',
                           dataset_class, instances))
    return out


def _from_via2(json_path, dataset_class, cwd):
    """VIA 2 project file -> polygon annotations; vertices move to pixel centres (+0.5)."""
    with open(json_path, 'rb') as f:
        project = json.load(f)
    picture_dir = Path(json_path.parent, project['_via_settings']['core']['default_filepath'])
    out = []
    for idx, entry in enumerate(project['_via_img_metadata'].values()):
        picture = Path(picture_dir, entry['filename'])
        attrs = entry['file_attributes']
        size = attrs.get('Size (width, height)', None)
        if size:
            width, height = (int(v) for v in size.split(', '))
        else:
            height, width = _imread(picture, as_gray=True).shape
        instances = []
        for region in entry['regions']:
            xs, ys = region['shape_attributes']['all_points_x'], region['shape_attributes']['all_points_y']
            ring = [v + 0.5 for xy in zip(xs, ys) for v in xy]
            instances.append(_instance(np.asarray((np.min(xs), np.min(ys), np.max(xs), np.max(ys))), [ring]))
        out.append(_record(idx, picture.relative_to(cwd), json_path.name, height, width, 'polygon', dataset_class,
                           instances, HFW=attrs.get('HFW', None)))
    return out


def _from_rle_json(json_path, dataset_class, cwd):
    """JSON list of {'file_name', 'segmentations': [COCO RLE with str counts]}; boxes straight from the runs."""
    with open(json_path, 'r') as f:
        listing = json.load(f)
    out = []
    for idx, item in enumerate(listing):
        masks = item['segmentations']
        for m in masks:
            m['counts'] = m['counts'].encode('utf-8')
        height, width = masks[0]['size']
        area, tight = engine.measure_rle(masks)
        instances = [_instance(tight[k].astype(np.float64) if area[k] else np.zeros(4), m) for k, m in enumerate(masks)]
        out.append(_record(idx, Path(json_path.parent, Path(item['file_name'])).relative_to(cwd), str(json_path),
                           height, width, 'bitmask', dataset_class, instances))
    return out


def get_ddicts(label_fmt, im_root, ann_root=None, pattern='*', dataset_class=None):
    """Images + ground-truth annotations -> detectron2-style data dicts (data_utils.py:313-530).

    label_fmt 'binary' / 'label': annotation images (or .npy) in *ann_root*; instances are the
    connected components of the binary image (8-connectivity, raster order, as
    skimage.measure.label) or the distinct non-zero label values in ascending order.
    'via2': *im_root* is a VIA 2 JSON project -> polygons.  'rle': *im_root* is a JSON list of
    {'file_name', 'segmentations'}.  All instances get category 0 (single-class, like the reference)."""
    cwd = Path()
    im_root = Path(im_root)
    fmt = label_fmt.lower()
    if fmt in ('binary', 'label'):
        return _from_annotation_images(label_fmt == 'binary', im_root, Path(ann_root) if ann_root else None, pattern,
                                       dataset_class, cwd)
    if fmt == 'via2':
        return _from_via2(im_root, dataset_class, cwd)
    if fmt == 'rle':
        return _from_rle_json(im_root, dataset_class, cwd)
    raise (ValueError("label_fmt must be 'binary','label', or 'via2'"))
