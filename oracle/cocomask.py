"""oracle/cocomask.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A stand-in for ``pycocotools.mask`` (2.0.4, the version AMPIS pins at
/root/reference/docker/env.yml:21) backed by ``oracle/libmaskapi_ref.so``
(``oracle/maskapi_ref.c``).  Function names, argument order, return dtypes
and array layouts follow the pycocotools Python wrapper as described in
SURVEY.md Appendix A.1, so the AMPIS loops in ``oracle/ampis_ref.py`` read
like the reference's own code.

PARITY STATUS: parity unpinned against a real pycocotools binary (none is
available offline); see the header of maskapi_ref.c for what is pinned.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmaskapi_ref.so")


def build(force=False):
    """Compile the C oracle with the committed Makefile (gcc -O2, scalar)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "maskapi_ref.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "clean", "all"])
    return _LIB_PATH


def _load():
    build()
    lib = C.CDLL(_LIB_PATH)
    u32p = C.POINTER(C.c_uint32)
    siz = C.c_ulong
    lib.ref_rle_fr_string.restype = siz
    lib.ref_rle_fr_string.argtypes = [C.c_char_p, siz, u32p]
    lib.ref_rle_to_string.restype = siz
    lib.ref_rle_to_string.argtypes = [u32p, siz, C.c_char_p]
    lib.ref_rle_area.restype = C.c_uint32
    lib.ref_rle_area.argtypes = [u32p, siz]
    lib.ref_rle_decode.restype = None
    lib.ref_rle_decode.argtypes = [u32p, siz, C.c_void_p]
    lib.ref_rle_encode.restype = siz
    lib.ref_rle_encode.argtypes = [C.c_void_p, siz, siz, u32p]
    lib.ref_rle_merge2.restype = siz
    lib.ref_rle_merge2.argtypes = [u32p, siz, u32p, siz, siz, siz, C.c_int, u32p]
    lib.ref_rle_to_bbox.restype = None
    lib.ref_rle_to_bbox.argtypes = [u32p, siz, C.c_uint32, C.c_uint32, C.POINTER(C.c_double)]
    lib.ref_rle_fr_poly.restype = siz
    lib.ref_rle_fr_poly.argtypes = [C.POINTER(C.c_double), siz, siz, siz, u32p, siz]
    lib.ref_rle_fr_bbox.restype = siz
    lib.ref_rle_fr_bbox.argtypes = [C.POINTER(C.c_double), siz, siz, u32p, siz]
    lib.ref_iou_strings.restype = None
    lib.ref_iou_strings.argtypes = [C.POINTER(C.c_char_p), siz, C.POINTER(C.c_char_p), siz,
                                    C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    lib.ref_merge2_area_strings.restype = C.c_uint32
    lib.ref_merge2_area_strings.argtypes = [C.c_char_p, C.c_char_p, siz, siz, C.c_int]
    lib.ref_area_strings.restype = None
    lib.ref_area_strings.argtypes = [C.POINTER(C.c_char_p), siz, C.c_void_p]
    return lib


_lib = _load()
_u32p = C.POINTER(C.c_uint32)


def _as_u32(a):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    return a, a.ctypes.data_as(_u32p)


def _counts_bytes(obj):
    c = obj['counts']
    return c if isinstance(c, bytes) else c.encode('ascii')


# ---- uncompressed helpers (not part of pycocotools' Python API) -----------

def counts_from_string(s):
    """rleFrString: compressed bytes -> uint32 counts."""
    out = np.empty(max(len(s), 1), np.uint32)
    m = _lib.ref_rle_fr_string(s, len(s), out.ctypes.data_as(_u32p))
    return out[:m].copy()


def string_from_counts(cnts):
    """rleToString: uint32 counts -> compressed bytes."""
    a, p = _as_u32(cnts)
    buf = C.create_string_buffer(6 * len(a) + 1)
    n = _lib.ref_rle_to_string(p, len(a), buf)
    return buf.raw[:n]


def to_bbox(rle):
    """rleToBbox of one compressed RLE dict -> float64[4] (x, y, w, h)."""
    a, p = _as_u32(counts_from_string(_counts_bytes(rle)))
    bb = np.zeros(4, np.float64)
    h, w = rle['size']
    _lib.ref_rle_to_bbox(p, len(a), h, w, bb.ctypes.data_as(C.POINTER(C.c_double)))
    return bb


# ---- pycocotools.mask API --------------------------------------------------

def encode(bimask):
    """encode(uint8 [h,w] or [h,w,n], Fortran order) -> dict or list of dicts."""
    single = bimask.ndim == 2
    arr = bimask[:, :, None] if single else bimask
    h, w, n = arr.shape
    out = []
    scratch = np.empty(h * w + 1, np.uint32)
    for i in range(n):
        m = np.ascontiguousarray(arr[:, :, i].T.astype(np.uint8))  # column-major bytes
        k = _lib.ref_rle_encode(m.ctypes.data, h, w, scratch.ctypes.data_as(_u32p))
        out.append({'size': [h, w], 'counts': string_from_counts(scratch[:k])})
    return out[0] if single else out


def decode(rleObjs):
    """decode(dict) -> uint8[h,w]; decode(list) -> uint8[h,w,n] (Fortran order)."""
    single = isinstance(rleObjs, dict)
    objs = [rleObjs] if single else rleObjs
    h, w = objs[0]['size']
    out = np.zeros((h, w, len(objs)), np.uint8, order='F')
    for i, o in enumerate(objs):
        a, p = _as_u32(counts_from_string(_counts_bytes(o)))
        buf = np.zeros(int(a.sum()), np.uint8)
        _lib.ref_rle_decode(p, len(a), buf.ctypes.data)
        out[:, :, i] = buf[:h * w].reshape((w, h)).T
    return out[:, :, 0] if single else out


def area(rleObjs):
    """area(dict) -> uint32 scalar; area(list) -> uint32[n]."""
    single = isinstance(rleObjs, dict)
    objs = [rleObjs] if single else rleObjs
    n = len(objs)
    out = np.zeros(n, np.uint32)
    arr = (C.c_char_p * max(n, 1))(*[_counts_bytes(o) for o in objs])
    _lib.ref_area_strings(arr, n, out.ctypes.data)
    return out[0] if single else out


def merge(rleObjs, intersect=False):
    """merge(list of dicts, intersect) -> one compressed RLE dict (two-mask folds)."""
    if len(rleObjs) == 0:
        return {'size': [0, 0], 'counts': b''}
    h, w = rleObjs[0]['size']
    acc = counts_from_string(_counts_bytes(rleObjs[0]))
    for o in rleObjs[1:]:
        if list(o['size']) != [h, w]:
            return {'size': [0, 0], 'counts': b''}
        b = counts_from_string(_counts_bytes(o))
        a_, pa = _as_u32(acc)
        b_, pb = _as_u32(b)
        out = np.empty(len(a_) + len(b_) + 1, np.uint32)
        m = _lib.ref_rle_merge2(pa, len(a_), pb, len(b_), h, w, int(bool(intersect)),
                                out.ctypes.data_as(_u32p))
        acc = out[:m].copy()
    return {'size': [h, w], 'counts': string_from_counts(acc)}


def merge_area(a, b, intersect=True):
    """area(merge([a, b], intersect)) in ONE C call (string round trip kept)."""
    h, w = a['size']
    return np.uint32(_lib.ref_merge2_area_strings(_counts_bytes(a), _counts_bytes(b), h, w,
                                                  int(bool(intersect))))


def iou(dt, gt, pyiscrowd):
    """iou(dt, gt, iscrowd) -> float64[len(dt), len(gt)] (or [] if either is empty)."""
    m, n = len(dt), len(gt)
    if m == 0 or n == 0:
        return []
    crowd = np.ascontiguousarray(np.asarray(pyiscrowd, dtype=np.uint8))
    assert crowd.shape[0] == n, 'iscrowd must have the same length as gt'
    h, w = dt[0]['size']
    for o in list(dt) + list(gt):
        if list(o['size']) != [h, w]:
            raise ValueError('oracle iou: mixed image sizes are not supported')
    o = np.zeros(m * n, np.float64)
    da = (C.c_char_p * m)(*[_counts_bytes(x) for x in dt])
    ga = (C.c_char_p * n)(*[_counts_bytes(x) for x in gt])
    _lib.ref_iou_strings(da, m, ga, n, h, w, crowd.ctypes.data, o.ctypes.data)
    return o.reshape((m, n), order='F')


def frPyObjects(pyobj, h, w):
    """frPyObjects(list of polygons / bboxes, h, w) -> list of RLE dicts
    (or a single polygon / bbox -> one dict)."""
    def poly(p):
        xy = np.ascontiguousarray(np.asarray(p, dtype=np.float64))
        k = len(xy) // 2
        cap = 4 * (2 * (w + 2)) + 16 + 8 * k
        while True:
            out = np.empty(cap, np.uint32)
            m = _lib.ref_rle_fr_poly(xy.ctypes.data_as(C.POINTER(C.c_double)), k, h, w,
                                     out.ctypes.data_as(_u32p), cap)
            if m != C.c_ulong(-1).value:
                break
            cap *= 4
        return {'size': [h, w], 'counts': string_from_counts(out[:m])}

    def bbox(b):
        bb = np.ascontiguousarray(np.asarray(b, dtype=np.float64))
        out = np.empty(4 * (2 * (w + 2)) + 64, np.uint32)
        m = _lib.ref_rle_fr_bbox(bb.ctypes.data_as(C.POINTER(C.c_double)), h, w,
                                 out.ctypes.data_as(_u32p), len(out))
        return {'size': [h, w], 'counts': string_from_counts(out[:m])}

    if type(pyobj) == np.ndarray:
        return [bbox(b) for b in pyobj]
    if type(pyobj) == list and len(pyobj[0]) == 4:
        return [bbox(b) for b in pyobj]
    if type(pyobj) == list and len(pyobj[0]) > 4:
        return [poly(p) for p in pyobj]
    if type(pyobj) == list and len(pyobj) == 4:
        return bbox(pyobj)
    if type(pyobj) == list and len(pyobj) > 4:
        return poly(pyobj)
    raise Exception('input type is not supported.')
