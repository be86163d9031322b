"""``extract_boxes`` -- drop-in for the one function of ``ampis.data_utils`` that sits on the
mask-measurement path (reference ampis/data_utils.py:180-252).  The loaders, the trainer hook
and the detectron2 glue of that module are outside this package's scope."""
import numpy as np

from . import engine


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
    area, bb = engine.bool_area_bbox(np.ascontiguousarray(masks != 0))
    ne = area > 0
    x1, y1, x2, y2 = bb[ne, 0], bb[ne, 1], bb[ne, 2], bb[ne, 3]
    if box_mode == 'detectron2':
        boxes[ne] = np.stack([x1, y1, x2, y2], axis=1)
    else:
        boxes[ne] = np.stack([y1, y2 + 1, x1, x2 + 1], axis=1)
    return boxes
