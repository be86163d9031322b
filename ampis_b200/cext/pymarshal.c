/* _pymarshal -- host marshaller of the drop-in API (CPython C API, no CUDA): walks the Python containers the
 * reference API receives -- lists of COCO RLE dicts {'size': [h, w], 'counts': bytes} -- and writes, for every mask,
 * the ADDRESS and LENGTH of its compressed string and its image size into caller-provided arrays.  Nothing is
 * copied or joined: the strings stay where Python keeps them and libampis_b200.so gathers them straight into its
 * pinned staging buffer (ampis_eval_images_host).  A Python loop over the dicts costs ~150 ns per mask
 * (m['counts'], len(), m['size']), this walk ~15 ns -- at 1,000 masks per image that is the difference between
 * 7 k and 60 k images per second before the GPU is even asked.
 *
 * Replaces the per-call work of pycocotools' _frString (called from every RLE.iou / merge / area of
 * analyze.py:108,158,315-321 and powder.py:82-83) on the HOST side only: type checks and pointer extraction.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static PyObject *s_counts, *s_size;

static int get_dim(PyObject *o, long *out)
{
    long v;
#if PY_VERSION_HEX >= 0x030c0000
    if (PyLong_CheckExact(o) && PyUnstable_Long_IsCompact((const PyLongObject *)o)) {      /* one digit: no call */
        *out = (long)PyUnstable_Long_CompactValue((const PyLongObject *)o);
        return 0;
    }
#endif
    if (PyLong_Check(o)) {
        v = PyLong_AsLong(o);
    } else {
        PyObject *i = PyNumber_Index(o);        /* numpy integers */
        if (!i) return -1;
        v = PyLong_AsLong(i);
        Py_DECREF(i);
    }
    if (v == -1 && PyErr_Occurred()) return -1;
    *out = v;
    return 0;
}

/* gather(list_of_mask_lists, ptr_u64, len_i32, hw_i32[, keep]) -> (n_masks, first_mixed_size_list or -1)
 * ptr_u64[k], len_i32[k]: address and length of the k-th mask's compressed counts (masks of list 0, then list 1, ...)
 * hw_i32[2 k], hw_i32[2 k + 1]: its image size.  str counts are accepted (their cached UTF-8 form is used).
 * The second result is the index of the first inner list whose masks do not all share one size.
 * keep: optional list; every counts object is appended to it (references that keep the strings alive). */
static PyObject *gather(PyObject *self, PyObject *args)
{
    PyObject *lists, *o_ptr, *o_len, *o_hw, *keep = NULL;
    if (!PyArg_ParseTuple(args, "OOOO|O", &lists, &o_ptr, &o_len, &o_hw, &keep)) return NULL;
    if (keep == Py_None) keep = NULL;
    if (keep && !PyList_Check(keep)) { PyErr_SetString(PyExc_TypeError, "gather: keep must be a list"); return NULL; }
    Py_buffer b_ptr, b_len, b_hw;
    if (PyObject_GetBuffer(o_ptr, &b_ptr, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) return NULL;
    if (PyObject_GetBuffer(o_len, &b_len, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) { PyBuffer_Release(&b_ptr); return NULL; }
    if (PyObject_GetBuffer(o_hw, &b_hw, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) < 0) {
        PyBuffer_Release(&b_ptr); PyBuffer_Release(&b_len); return NULL;
    }
    uint64_t *ptr = (uint64_t *)b_ptr.buf;
    int32_t *len = (int32_t *)b_len.buf, *hw = (int32_t *)b_hw.buf;
    const Py_ssize_t cap = b_ptr.len / 8;
    PyObject *outer = NULL, *result = NULL;
    Py_ssize_t k = 0, mixed = -1;
    PyObject *last_sz = NULL, *last_i0 = NULL, *last_i1 = NULL;      /* borrowed, only compared */
    long last_h = 0, last_w = 0, items_h = 0, items_w = 0;
    if (b_len.len / 4 < cap || b_hw.len / 8 < cap) {
        PyErr_SetString(PyExc_ValueError, "gather: output arrays of different capacity");
        goto done;
    }
    outer = PySequence_Fast(lists, "gather: expected a sequence of mask lists");
    if (!outer) goto done;
    for (Py_ssize_t li = 0; li < PySequence_Fast_GET_SIZE(outer); li++) {
        PyObject *inner = PySequence_Fast(PySequence_Fast_GET_ITEM(outer, li), "gather: expected a list of RLE dicts");
        if (!inner) goto done;
        const Py_ssize_t n = PySequence_Fast_GET_SIZE(inner);
        const Py_ssize_t k0 = k;
        PyObject **items = PySequence_Fast_ITEMS(inner);
        for (Py_ssize_t i = 0; i < n; i++, k++) {
            PyObject *m = items[i];
            /* the walk is a chain of cache misses (dict -> its table -> the bytes object / the size list); ask for
             * the objects a few masks ahead so that the misses of neighbouring masks overlap */
            if (i + 8 < n) __builtin_prefetch(items[i + 8]);
            if (i + 4 < n && PyDict_Check(items[i + 4])) __builtin_prefetch(((PyDictObject *)items[i + 4])->ma_keys);
            if (k >= cap) { PyErr_SetString(PyExc_ValueError, "gather: more masks than capacity"); Py_DECREF(inner); goto done; }
            if (!PyDict_Check(m)) {
                PyErr_Format(PyExc_TypeError, "RLE masks must be dicts with 'size' and 'counts', got %s", Py_TYPE(m)->tp_name);
                Py_DECREF(inner); goto done;
            }
            PyObject *c = PyDict_GetItemWithError(m, s_counts), *sz = PyDict_GetItemWithError(m, s_size);
            if (!c || !sz) {
                if (!PyErr_Occurred()) PyErr_SetString(PyExc_KeyError, c ? "size" : "counts");
                Py_DECREF(inner); goto done;
            }
            if (PyBytes_Check(c)) {
                ptr[k] = (uint64_t)(uintptr_t)PyBytes_AS_STRING(c);
                len[k] = (int32_t)PyBytes_GET_SIZE(c);
            } else if (PyUnicode_Check(c)) {
                Py_ssize_t sl;
                const char *s = PyUnicode_AsUTF8AndSize(c, &sl);
                if (!s) { Py_DECREF(inner); goto done; }
                ptr[k] = (uint64_t)(uintptr_t)s;
                len[k] = (int32_t)sl;
            } else if (PyByteArray_Check(c)) {
                ptr[k] = (uint64_t)(uintptr_t)PyByteArray_AS_STRING(c);
                len[k] = (int32_t)PyByteArray_GET_SIZE(c);
            } else {
                PyErr_Format(PyExc_TypeError, "RLE counts must be compressed bytes/str, got %s", Py_TYPE(c)->tp_name);
                Py_DECREF(inner); goto done;
            }
            /* the caller's cache keeps the string objects alive: their addresses are its key */
            if (keep && PyList_Append(keep, c) < 0) { Py_DECREF(inner); goto done; }
            long hh, ww;
            if (sz == last_sz) {                          /* the same size object as the mask before: parsed already */
                hh = last_h; ww = last_w;
            } else if (PyList_CheckExact(sz) && PyList_GET_SIZE(sz) >= 2) {
                PyObject *i0 = PyList_GET_ITEM(sz, 0), *i1 = PyList_GET_ITEM(sz, 1);
                if (i0 == last_i0 && i1 == last_i1) {     /* [h, w] built from the same two int objects: same values */
                    hh = items_h; ww = items_w;
                } else {
                    if (get_dim(i0, &hh) < 0 || get_dim(i1, &ww) < 0) { Py_DECREF(inner); goto done; }
                    last_i0 = i0; last_i1 = i1; items_h = hh; items_w = ww;
                }
            } else {
                PyObject *f = PySequence_Fast(sz, "RLE 'size' must be a pair [h, w]");
                if (!f) { Py_DECREF(inner); goto done; }
                if (PySequence_Fast_GET_SIZE(f) < 2) {
                    PyErr_SetString(PyExc_ValueError, "RLE 'size' must be a pair [h, w]");
                    Py_DECREF(f); Py_DECREF(inner); goto done;
                }
                const int bad = get_dim(PySequence_Fast_GET_ITEM(f, 0), &hh) < 0 || get_dim(PySequence_Fast_GET_ITEM(f, 1), &ww) < 0;
                Py_DECREF(f);
                if (bad) { Py_DECREF(inner); goto done; }
            }
            if (hh < 0 || ww < 0 || hh > 0x7fffffffL || ww > 0x7fffffffL) {
                PyErr_SetString(PyExc_ValueError, "RLE 'size' out of range");
                Py_DECREF(inner); goto done;
            }
            last_sz = sz; last_h = hh; last_w = ww;
            hw[2 * k] = (int32_t)hh;
            hw[2 * k + 1] = (int32_t)ww;
            if (mixed < 0 && k > k0 && (hw[2 * k] != hw[2 * k0] || hw[2 * k + 1] != hw[2 * k0 + 1])) mixed = li;
        }
        Py_DECREF(inner);
    }
    result = Py_BuildValue("(nn)", k, mixed);
done:
    Py_XDECREF(outer);
    PyBuffer_Release(&b_ptr); PyBuffer_Release(&b_len); PyBuffer_Release(&b_hw);
    return result;
}

static PyMethodDef methods[] = {
    {"gather", gather, METH_VARARGS, "gather(list_of_mask_lists, ptr_u64, len_i32, hw_i32[, keep]) -> (n_masks, first_mixed_list)"},
    {NULL, NULL, 0, NULL}
};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_pymarshal", "host marshaller of ampis_b200", -1, methods};

PyMODINIT_FUNC PyInit__pymarshal(void)
{
    s_counts = PyUnicode_InternFromString("counts");
    s_size = PyUnicode_InternFromString("size");
    if (!s_counts || !s_size) return NULL;
    return PyModule_Create(&module);
}
