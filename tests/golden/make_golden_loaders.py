#!/usr/bin/env python
"""Golden fixtures for the annotation-image loaders (run in the build container, where
/root/reference is mounted; the GPU box only sees the committed .npz):

    python tests/golden/make_golden_loaders.py

Stores two of the reference's binary annotation PNGs (examples/spheroidite/data/annotations, bit
packed) and what the oracle restatement of data_utils.get_ddicts('binary' / 'label') yields for
them: number of instances, boxes, RLE strings."""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ampis_ref as R  # noqa: E402

REF = '/root/reference/examples/spheroidite/data/annotations'
NAMES = ['train_800C-24H-Q-2_sizeRC_484_645.png', 'valid_800C-85H-Q-6_sizeRC_483_645.png']


def main():
    out = {'names': np.array(NAMES)}
    for k, name in enumerate(NAMES):
        a = np.asarray(Image.open(os.path.join(REF, name)))
        out['%d_shape' % k] = np.array(a.shape)
        out['%d_values' % k] = np.unique(a)
        out['%d_bits' % k] = np.packbits(a.astype(bool).ravel())
        anns = R.annotations_from_label_image(a, binary=True)
        out['%d_boxes' % k] = np.stack([b for b, _ in anns])
        strings = [m['counts'] for _, m in anns]
        off = np.zeros(len(strings) + 1, np.int64)
        np.cumsum([len(s) for s in strings], out=off[1:])
        out['%d_blob' % k] = np.frombuffer(b''.join(strings), np.uint8)
        out['%d_off' % k] = off
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'spheroidite_annotations.npz'), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
