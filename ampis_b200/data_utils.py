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


def get_ddicts(label_fmt, im_root, ann_root=None, pattern='*', dataset_class=None):
    """Images + ground-truth annotations -> detectron2-style data dicts (data_utils.py:313-530).

    label_fmt 'binary' / 'label': annotation images (or .npy) in *ann_root*; instances are the
    connected components of the binary image (8-connectivity, raster order, as
    skimage.measure.label) or the distinct non-zero label values in ascending order.  Labelling,
    boxes and RLE encoding run on the GPU (csrc/label.cu) from ONE pass over the image instead of a
    full-frame pass per instance.  'via2': VIA 2 JSON -> polygons (+0.5 pixel-centre shift).
    'rle': JSON list of {'file_name', 'segmentations'}; boxes come from the runs
    (csrc/rle_measure.cu) without decoding."""
    cwd = Path()
    im_root = Path(im_root)
    ann_root = Path(ann_root) if ann_root else None
    ddicts = []

    if label_fmt.lower() in ['binary', 'label']:
        img_paths = Path(im_root).glob(pattern)
        for idx, p in enumerate(img_paths):
            file_annotations = list(Path(ann_root).glob('*{}*'.format(p.stem)))
            n = len(file_annotations)
            assert n == 1, f'There must be exactly 1 annotation file for, {p.name}, but {n} were found'
            ann_path = file_annotations[0].relative_to(cwd)
            ann = np.load(str(ann_path)) if ann_path.suffix == '.npy' else _imread(ann_path)
            height, width = ann.shape[:2]
            ddict = {'file_name': str(p.relative_to(cwd)),
                     'annotation_file': str(ann_path),
                     'height': height,
                     'width': width,
                     'mask_format': 'bitmask',
                     'image_id': idx,
                     'dataset_class': dataset_class}
            rles, bb = engine.label_image_to_instances(ann, binary=(label_fmt == 'binary'))
            annotations = []
            for mask, b in zip(rles, bb):
                annotations.append({'bbox': b.astype(np.float64),          # extract_boxes(mask)[0]: x1, y1, x2, y2
                                    'bbox_mode': BoxMode.XYXY_ABS,
                                    'segmentation': mask,
                                    'category_id': 0})
            ddict['annotations'] = annotations
            ddict['num_instances'] = len(annotations)
            ddicts.append(ddict)

    elif label_fmt.lower() == 'via2':
        with open(Path(im_root), 'rb') as f:
            j = json.load(f)
        img_dir = Path(im_root.parent, j['_via_settings']['core']['default_filepath'])
        for idx, annos in enumerate(j['_via_img_metadata'].values()):
            filename = Path(img_dir, annos['filename'])
            size = annos['file_attributes'].get('Size (width, height)', None)
            if size:
                width, height = tuple((int(x) for x in size.split(', ')))
            else:
                im = _imread(filename, as_gray=True)
                height, width = im.shape
            hfw = annos['file_attributes'].get('HFW', None)
            ddict = {'file_name': str(filename.relative_to(cwd)),
                     'annotation_file': im_root.name,
                     'height': height,
                     'width': width,
                     'mask_format': 'polygon',
                     'image_id': idx,
                     'HFW': hfw,
                     'dataset_class': dataset_class}
            annotations = []
            for obj in annos['regions']:
                shape = obj['shape_attributes']
                px = shape['all_points_x']
                py = shape['all_points_y']
                poly = [(x + 0.5, y + 0.5) for x, y in zip(px, py)]
                poly = [p for x in poly for p in x]
                annotations.append({'bbox': np.asarray((np.min(px), np.min(py), np.max(px), np.max(py))),
                                    'bbox_mode': BoxMode.XYXY_ABS,
                                    'segmentation': [poly],
                                    'category_id': 0})
            ddict['annotations'] = annotations
            ddict['num_instances'] = len(annotations)
            ddicts.append(ddict)

    elif label_fmt.lower() == 'rle':
        with open(im_root, 'r') as f:
            data = json.load(f)
        for i, anns in enumerate(data):
            for jj, ann in enumerate(anns['segmentations']):
                data[i]['segmentations'][jj]['counts'] = ann['counts'].encode('utf-8')
        for idx, p in enumerate(data):
            img_path = Path(im_root.parent, Path(p['file_name']))
            ann = p['segmentations']
            height, width = ann[0]['size']
            ddict = {'file_name': str(img_path.relative_to(cwd)),
                     'annotation_file': str(im_root),
                     'height': height,
                     'width': width,
                     'mask_format': 'bitmask',
                     'image_id': idx,
                     'dataset_class': dataset_class}
            table = engine.table_from_rle(ann, paint=False)          # boxes straight from the runs
            area, bb = table.areas_np(), table.bbox_np()
            annotations = []
            for k, mask in enumerate(ann):
                bbox = bb[k].astype(np.float64) if area[k] else np.zeros(4)
                annotations.append({'bbox': bbox,
                                    'bbox_mode': BoxMode.XYXY_ABS,
                                    'segmentation': mask,
                                    'category_id': 0})
            ddict['annotations'] = annotations
            ddict['num_instances'] = len(annotations)
            ddicts.append(ddict)
    else:
        raise (ValueError("label_fmt must be 'binary','label', or 'via2'"))
    return ddicts
