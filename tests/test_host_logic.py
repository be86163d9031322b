"""CPU tests of the host side: the C ABI library loads and exports every declared symbol,
containers behave like the reference's, group bookkeeping and the synthetic generator are
consistent.  No kernel is launched here."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ampis_b200 import _native
    lib = _native.lib()
    hdr = open(os.path.join(ROOT, 'include', 'ampis_b200.h')).read()
    declared = set(re.findall(r'\b(ampis_[a-z0-9_]+)\s*\(', hdr))
    assert declared, 'no declarations found'
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ampis_version() >= 100


def test_no_cpu_fallback_without_device():
    import torch
    from ampis_b200 import analyze, engine
    if torch.cuda.is_available():
        pytest.skip('a device is present')
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        engine.require_cuda()
    m = {'size': [4, 4], 'counts': b'02208'}
    with pytest.raises(RuntimeError):
        analyze._piecewise_rle_match([m], [m])


def test_product_never_imports_oracle():
    import glob
    for f in glob.glob(os.path.join(ROOT, 'ampis_b200', '**', '*.py'), recursive=True):
        src = open(f).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, re.M), f
        assert 'oracle.' not in src and 'libmaskapi_ref' not in src, f


def test_rlemasks_indexing():
    import torch
    from ampis_b200.structures import RLEMasks
    r = RLEMasks([{'size': [2, 2], 'counts': bytes([48 + i])} for i in range(5)])
    assert len(r) == 5
    assert type(r[1].rle) == dict and len(r[1]) == 2               # quirk B.13: int wraps a dict
    assert [m['counts'] for m in r[1:3].rle] == [b'1', b'2']
    assert [m['counts'] for m in r[[4, 0]].rle] == [b'4', b'0']
    assert [m['counts'] for m in r[np.array([3, 1])].rle] == [b'3', b'1']
    mask = np.array([True, False, True, False, False])
    assert [m['counts'] for m in r[mask].rle] == [b'0', b'2']
    assert [m['counts'] for m in r[[bool(x) for x in mask]].rle] == [b'0', b'2']
    assert [m['counts'] for m in r[torch.tensor(mask)].rle] == [b'0', b'2']
    with pytest.raises(AssertionError):
        r[np.array([True, False])]


def test_instances_container_and_instance_set():
    from ampis_b200.containers import Instances, PolygonMasks, load_pickle  # noqa: F401
    from ampis_b200.structures import InstanceSet, RLEMasks
    masks = RLEMasks([{'size': [4, 4], 'counts': b'02208'}] * 3)
    inst = Instances((4, 4), masks=masks, boxes=np.zeros((3, 4)), class_idx=np.arange(3))
    assert len(inst) == 3 and inst.image_size == (4, 4) and inst.has('boxes') and not inst.has('scores')
    sub = inst[np.array([True, False, True])]
    assert len(sub) == 2 and sub.class_idx.tolist() == [0, 2] and len(sub.masks) == 2
    with pytest.raises(AssertionError):
        inst.scores = np.zeros(2)
    ddict = {'file_name': 'a/b.png', 'mask_format': 'polygon', 'height': 10, 'width': 12, 'HFW': '103.6 um',
             'annotations': [{'category_id': 0, 'bbox': [1, 1, 5, 5],
                              'segmentation': [[1.5, 1.5, 5.5, 1.5, 5.5, 5.5, 1.5, 5.5]]}]}
    iset = InstanceSet().read_from_ddict(ddict, inplace=False)
    assert iset.HFW == 103.6 and iset.HFW_units == 'um' and iset.pred_or_gt == 'gt'
    assert type(iset.instances.masks) == PolygonMasks and iset.instances.colors.shape == (1, 3)
    outs = {'file_name': 'x.png', 'dataset': 'powder_Training',
            'pred': {'instances': Instances((4, 4), pred_masks=masks.rle, pred_boxes=np.zeros((3, 4)),
                                            pred_classes=np.zeros(3, int), scores=np.ones(3))}}
    pset = InstanceSet().read_from_model_out(outs, inplace=False)
    assert pset.dataset_class == 'Training' and type(pset.instances.masks) == RLEMasks and len(pset.instances) == 3
    c = pset.copy()
    assert c is not pset and len(c.instances) == 3


def test_align_and_merge_boxes_and_shoelace():
    from ampis_b200 import analyze, structures

    class F:
        def __init__(self, p):
            self.filepath = p
    a = [F('x/1.png'), F('x/2.png'), F('x/3.png')]
    b = [F('y/3.png'), F('y/1.png')]
    ao, bo = analyze.align_instance_sets(a, b)
    assert [i.filepath for i in ao] == ['x/1.png', 'x/3.png'] and [i.filepath for i in bo] == ['y/1.png', 'y/3.png']
    assert analyze.merge_boxes([1, 5, 2, 6], [0, 4, 3, 9]).tolist() == [0, 5, 2, 9]
    assert structures._shoelace_area(np.array([0., 4, 4, 0]), np.array([0., 0, 3, 3])) == 12.0
    assert structures.boxes_to_array([[1, 2, 3, 4]]).shape == (1, 4)
    assert analyze.fast_instance_match is analyze.rle_instance_matcher


def test_group_bookkeeping_host_side():
    import torch
    from ampis_b200 import engine
    g = engine.Groups.interleaved(torch.device('cpu'), [2, 0, 3], [4, 5, 1], dense=True)
    assert g.row_mask.tolist() == [0, 1, 11, 12, 13] and g.row_grp.tolist() == [0, 0, 2, 2, 2]
    assert g.grp_row_begin.tolist() == [0, 2, 2] and g.grp_col_begin.tolist() == [2, 6, 14]
    assert g.grp_col_count.tolist() == [4, 5, 1] and g.max_cols == 5
    assert g.h_imat_off.tolist() == [0, 8, 8, 11] and g.imat_size == 11


def test_rle_marshalling_host_side():
    """Input handling of the RLE dict lists (pycocotools' _frString accepts bytes and str), size checks, the
    staging upload helper and the operand-fill arithmetic of the one-call result -- no device needed."""
    import torch
    from ampis_b200 import analyze, engine
    masks = [{'size': [6, 5], 'counts': b'02208'}, {'size': (6, 5), 'counts': '5d0'},
             {'size': np.array([6, 5]), 'counts': bytearray(b'0n0')}]
    strings, hs, ws = engine._rle_fields(masks)
    assert strings == [b'02208', b'5d0', b'0n0'] and all(type(x) is bytes for x in strings)
    assert hs.tolist() == [6, 6, 6] and ws.tolist() == [5, 5, 5]
    with pytest.raises(TypeError):
        engine._rle_fields([{'size': [6, 5], 'counts': [1, 2, 3]}])
    analyze._check_same_size(masks[:2], masks[2:])
    analyze._check_same_size([], [])
    with pytest.raises(ValueError, match=r'different image sizes.*\(6, 5\) vs \(5, 6\)'):
        analyze._check_same_size(masks, [{'size': [5, 6], 'counts': b'0n0'}])
    a, b, c = engine._upload(torch.device('cpu'), np.arange(5, dtype=np.int64), np.zeros(0, np.int32),
                             np.array([7, 8, 9], np.uint8))
    assert a.tolist() == [0, 1, 2, 3, 4] and a.dtype == torch.int64 and b.numel() == 0 and c.tolist() == [7, 8, 9]
    r = engine.ImageRows()
    r.hw, r.area = (256, 256), np.zeros(4, np.uint32)                       # 512 slabs of 128 pixels per mask
    r.span = np.array([[0, 512], [100, 356], [0, 0], [511, 512]], np.uint32)
    assert r.fill() == (512 + 256 + 0 + 1) / (4 * 512)


def test_c_marshaller_walks_the_dict_lists():
    """ampis_b200/cext/pymarshal.c (CPython C API, no device): address and length of every compressed string and the
    image size of every mask, for all the container shapes the reference API is handed -- bytes / str / bytearray
    counts; a size list per dict, ONE list shared by all dicts, lists built from the same two int objects (the three
    fast paths), tuples and numpy integers (the generic path) -- plus the first list whose masks differ in size, the
    references that keep the strings alive for the result cache, and the errors."""
    import ctypes
    from ampis_b200 import engine
    m = engine.marshal()
    rng = np.random.default_rng(11)

    def walk(lists, keep=None, cap=None):
        n = sum(len(x) for x in lists)
        cap = n if cap is None else cap
        ptr, ln, hw = np.zeros(max(cap, 1), np.uint64), np.zeros(max(cap, 1), np.int32), np.zeros((max(cap, 1), 2), np.int32)
        k, mixed = m.gather(lists, ptr, ln, hw) if keep is None else m.gather(lists, ptr, ln, hw, keep)
        assert k == n
        got = [ctypes.string_at(int(ptr[i]), int(ln[i])) for i in range(n)]
        return got, hw[:n].tolist(), mixed

    def counts(i):
        return bytes(rng.integers(48, 112, 3 + i % 9, dtype=np.uint8))

    shared, h, w = [37, 41], 37, 41
    big = 100000                                            # not a cached small int: every int(...) is a new object
    for size_of in (lambda i: [37, 41], lambda i: shared, lambda i: [h, w], lambda i: (37, 41),
                    lambda i: np.array([37, 41]), lambda i: [np.int32(37), np.int64(41)],
                    lambda i: [int(str(big)), int(str(big + 1))]):
        lists = [[{'size': size_of(i), 'counts': counts(i)} for i in range(n)] for n in (3, 0, 5, 1)]
        got, hw, mixed = walk(lists)
        want_hw = list(size_of(0)) if not isinstance(size_of(0), np.ndarray) else size_of(0).tolist()
        assert got == [d['counts'] for l in lists for d in l] and mixed == -1
        assert hw == [[int(want_hw[0]), int(want_hw[1])]] * 9
    # str and bytearray counts; the kept references
    lists = [[{'size': [5, 6], 'counts': 'abc0'}, {'size': [5, 6], 'counts': bytearray(b'0n0')}, {'size': [5, 6], 'counts': b'5d0'}]]
    keep = []
    got, hw, mixed = walk(lists, keep)
    assert got == [b'abc0', b'0n0', b'5d0'] and len(keep) == 3 and keep[0] is lists[0][0]['counts']
    # sizes: the first list whose masks differ (each way the fast paths could hide it)
    a, b = [7, 9], [7, 9]
    for odd in ([9, 7], (7, 10), [7, 9, 3][:2][::-1]):
        lists = [[{'size': a, 'counts': b'0'}, {'size': b, 'counts': b'1'}],
                 [{'size': a, 'counts': b'2'}, {'size': odd, 'counts': b'3'}, {'size': a, 'counts': b'4'}],
                 [{'size': odd, 'counts': b'5'}]]
        got, hw, mixed = walk(lists)
        assert mixed == 1 and hw[3] == [int(odd[0]), int(odd[1])] and hw[4] == [7, 9] and hw[5] == hw[3]
    # a generic-path size between two fast-path ones must not leave stale cached values behind
    i0, i1 = 300, 400
    lists = [[{'size': [i0, i1], 'counts': b'0'}, {'size': (500, 600), 'counts': b'1'}, {'size': [i0, i1], 'counts': b'2'}]]
    assert walk(lists)[1] == [[300, 400], [500, 600], [300, 400]]
    # errors
    for bad, exc in (([[{'size': [1, 2]}]], KeyError), ([[{'counts': b'0'}]], KeyError), ([[[1, 2]]], TypeError),
                     ([[{'size': [1, 2], 'counts': [1, 2]}]], TypeError), ([[{'size': [1], 'counts': b'0'}]], ValueError),
                     ([[{'size': [-1, 2], 'counts': b'0'}]], ValueError), ([[{'size': ['a', 2], 'counts': b'0'}]], TypeError),
                     ([3], TypeError)):
        with pytest.raises(exc):
            walk(bad, cap=4)
    with pytest.raises(ValueError, match='more masks than capacity'):
        walk([[{'size': [1, 2], 'counts': b'0'}] * 3], cap=2)


def test_eval_images_wrapper_hands_the_library_the_right_arrays(monkeypatch):
    """engine.eval_images without a device: the library call is replaced by a stand-in that READS its inputs and WRITES
    its outputs through the raw addresses it is given, exactly as libampis_b200.so would.  Checks the marshalled string
    addresses / lengths, the per-image counts and sizes, that every output array comes back filled (addresses of
    empty and non-empty arrays alike), and the per-image bookkeeping of det_seg_scores on top -- for one image (the
    plain-Python fast path) and for several."""
    import ctypes as C
    import torch
    from ampis_b200 import analyze, engine
    rng = np.random.default_rng(0)
    seen = {}

    def mk(n, size=(768, 1024)):
        return [{'size': list(size), 'counts': bytes(rng.integers(48, 112, int(rng.integers(60, 110)), dtype=np.uint8))}
                for _ in range(n)]

    def arr(addr, n, ct):
        return np.ctypeslib.as_array(C.cast(addr, C.POINTER(ct)), shape=(max(n, 1),))[:n]

    class StandIn(object):
        def ampis_eval_images_host(self, ptr, ln, n_img, n_rows, n_cols, h32, w32, mode, flags, crowd, d_ws, d_n, h_ws,
                                   h_n, best_col, best_inter, best_score, area, bbox, span, status, thr, n_thr, counts,
                                   totals, found, crowded, need, stream):
            G, P = arr(n_rows, n_img, C.c_int32).copy(), arr(n_cols, n_img, C.c_int32).copy()
            n, R = int((G + P).sum()), int(G.sum())
            p_, l_ = arr(ptr, n, C.c_uint64), arr(ln, n, C.c_int32)
            seen.update(strings=[C.string_at(int(p_[i]), int(l_[i])) for i in range(n)], G=G.tolist(), P=P.tolist(),
                        hw=(arr(h32, n_img, C.c_uint32).tolist(), arr(w32, n_img, C.c_uint32).tolist()))
            bc, bs, bi = arr(best_col, R, C.c_int32), arr(best_score, R, C.c_double), arr(best_inter, R, C.c_uint32)
            r0 = 0
            for g in range(n_img):
                for k in range(G[g]):
                    bc[r0 + k] = k % P[g] if P[g] else -1
                    bs[r0 + k] = 0.9 if (k % 3 and P[g]) else 0.0
                    bi[r0 + k] = 7
                r0 += G[g]
            arr(area, n, C.c_uint32)[:] = 11
            arr(status, n, C.c_int32)[:] = 0
            arr(bbox, 4 * n, C.c_int32)[:] = 1
            arr(span, 2 * n, C.c_uint32)[:] = 2
            return 0

    ws = [torch.empty(1 << 16, dtype=torch.uint8), torch.empty(1 << 16, dtype=torch.uint8)]
    monkeypatch.setattr(engine, 'require_cuda', lambda: torch.device('cpu'))
    monkeypatch.setattr(engine, '_workspaces', lambda device: ws)
    monkeypatch.setattr(engine, '_stream', lambda: None)
    monkeypatch.setattr(engine.N, 'lib', lambda: StandIn())
    for G, P in ((5, 7), (1, 1), (0, 3), (4, 0), (300, 300)):
        gt, pr = mk(G), mk(P)
        r = engine.eval_images([gt], [pr], engine.MODE_IOU)
        assert seen['strings'] == [m['counts'] for m in gt + pr] and seen['G'] == [G] and seen['P'] == [P]
        assert seen['hw'] == ([768], [1024]) and r.hw.tolist() == [[768, 1024]]
        assert (r.area == 11).all() and (r.status == 0).all() and len(r.best_col) == G and len(r.area) == G + P
        if G and P:
            assert r.best_col.tolist() == [k % P for k in range(G)] and (r.bbox == 1).all() and (r.span == 2).all()
            d = analyze.det_seg_scores(gt, pr, 0.5)
            assert len(d['det_tp']) == sum(1 for k in range(G) if k % 3) and (d['seg_tp'] == 7).all()
            if len(d['det_tp']):
                assert d['det_tp'][:, 0].tolist() == [k for k in range(G) if k % 3]
    gts, prs = [mk(3), mk(0), mk(2)], [mk(2), mk(4), mk(5, (64, 48))]
    with pytest.raises(ValueError, match=r'different image sizes.*\(768, 1024\) vs \(64, 48\)'):
        engine.eval_images(gts, prs, engine.MODE_IOU)
    prs[2] = mk(5)
    r = engine.eval_images(gts, prs, engine.MODE_IOU)
    assert seen['strings'] == [m['counts'] for g, p in zip(gts, prs) for m in g + p]
    assert seen['G'] == [3, 0, 2] and seen['P'] == [2, 4, 5] and r.row_off.tolist() == [0, 3, 3, 5]
    with pytest.raises(ValueError, match='different image sizes'):
        engine.eval_images([mk(2) + mk(1, (5, 6))], [mk(2)], engine.MODE_IOU)
    # powder's satellite matcher leaves the size check to the marshaller and keeps its own message
    from ampis_b200.applications import powder
    with pytest.raises(ValueError, match='particle and satellite masks must share one image size'):
        powder._rle_satellite_match(mk(4), mk(2) + mk(1, (5, 6)))
    with pytest.raises(ValueError, match='particle and satellite masks must share one image size'):
        powder._rle_satellite_match(mk(4, (5, 6)), mk(3))
    out = powder._rle_satellite_match(mk(4), mk(6))          # the stand-in scores 0.9 for rows 1, 2, 4, 5
    assert out['satellite_matches'].tolist() == [[1, 1], [2, 2], [4, 0], [5, 1]] and out['satellites_unmatched'].tolist() == [0, 3]


def test_model_of_the_warp_string_decoder_equals_the_oracle():
    """Lane-by-lane model (plain Python) of rle_string_decode_warp_kernel (csrc/rle_codec.cu): four characters per lane
    read as aligned words, numbers that end inside a lane, the start of a lane's first number from the nearest lower
    lane with a number end and the two lanes after it (sd_incoming), compaction, the two delta chains and the carries
    between steps of 128 characters -- against the oracle's rleFrString on valid encodings and on arbitrary numbers of
    1 to 10 characters, at every byte alignment.  The GPU test of the same name runs the kernel itself."""
    from oracle import cocomask as rle
    M32 = 0xffffffff
    shl = lambda x, sft: (x << sft) & M32 if sft < 32 else 0

    def incoming(E, L, pk, cx, ck):
        below = E if L >= 32 else (E & ((1 << L) - 1))
        pe = below.bit_length() - 1 if below else -1
        g0, g1, g2 = pk[max(pe, 0)], pk[min(pe + 1, 31)], pk[min(pe + 2, 31)]
        pv, pl = (g0 & 0xffffff, g0 >> 24) if pe >= 0 else (cx, ck)
        if pe + 1 < L:
            pv |= shl(g1 & 0xffffff, 5 * pl); pl += g1 >> 24
        if pe + 2 < L:
            pv |= shl(g2 & 0xffffff, 5 * pl); pl += g2 >> 24
        if pe + 3 < L:
            pl = 8
        return pv, min(pl, 8)

    def decode(buf, start, ln):
        a, out, m_base, cx, ck, ce, co, p0 = start & 3, {}, 0, 0, 0, 0, 0, 0
        while p0 < ln:
            nloc, v, hk, hs, pk = [0] * 32, [[0] * 4 for _ in range(32)], [0] * 32, [False] * 32, [0] * 32
            for lane in range(32):
                rem = ln - (p0 + 4 * lane)
                nv = 0 if rem <= 0 else min(rem, 4)
                lastb = nv - 1 if 1 <= rem <= 4 else 8
                byts = 0
                if nv:
                    wi = (start - a) + 4 * ((p0 >> 2) + lane)
                    w0 = int.from_bytes(buf[wi:wi + 4].ljust(4, b'\0'), 'little')
                    w1 = int.from_bytes(buf[wi + 4:wi + 8].ljust(4, b'\0'), 'little') if a + nv > 4 else 0
                    byts = ((w1 << 32 | w0) >> (8 * a)) & M32
                cur = curk = 0
                for b in range(nv):
                    c = (((byts >> (8 * b)) & 0xff) - 48) & M32
                    cur |= (c & 0x1f) << (5 * curk); curk += 1
                    if not (c & 0x20) or b == lastb:
                        sign = (not (c & 0x20)) and bool(c & 0x10)
                        x = cur
                        if nloc[lane] == 0:
                            hk[lane], hs[lane] = curk, sign
                        elif sign:
                            x |= (M32 << (5 * curk)) & M32
                        v[lane][nloc[lane]] = x; nloc[lane] += 1; cur = curk = 0
                pk[lane] = cur | (curk << 24)
            E = sum(1 << l for l in range(32) if nloc[l])
            S, slot = [0] * 132, 0
            for lane in range(32):
                pv, pl = incoming(E, lane, pk, cx, ck)
                if nloc[lane]:
                    v[lane][0] = pv | shl(v[lane][0], 5 * pl)
                    if hs[lane]:
                        v[lane][0] |= shl(M32, 5 * (pl + hk[lane]))
                S[slot:slot + nloc[lane]] = v[lane][:nloc[lane]]
                slot += nloc[lane]
            nn, odd = slot, m_base & 1
            ys = [[(S[4 * l + q] if 4 * l + q < nn and m_base + 4 * l + q != 0 else 0) for q in range(4)] for l in range(32)]
            A, B = [(y[0] + y[2]) & M32 for y in ys], [(y[1] + y[3]) & M32 for y in ys]
            SA, SB = np.cumsum(A) & M32, np.cumsum(B) & M32
            cA, cB = (co, ce) if odd else (ce, co)
            for l in range(32):
                y = ys[l]
                w0 = (int(SA[l]) - A[l] + y[0] + cA) & M32; w1 = (int(SB[l]) - B[l] + y[1] + cB) & M32
                w = [w0, w1, (w0 + y[2]) & M32, (w1 + y[3]) & M32]
                for q in range(4):
                    if 4 * l + q < nn:
                        out[m_base + 4 * l + q] = S[4 * l + q] if m_base + 4 * l + q == 0 else w[q]
            if p0 + 128 < ln:
                nA, nB = (cA + int(SA[31])) & M32, (cB + int(SB[31])) & M32
                ce, co = (nB, nA) if odd else (nA, nB)
                cx, ck = incoming(E, 32, pk, cx, ck)
            m_base += nn
            p0 += 128
        return np.array([out[k] for k in range(m_base)], np.uint32)

    rng = np.random.default_rng(5)
    for t in range(160):
        strs = []
        for _ in range(3):
            if rng.random() < 0.5:
                m, mag = int(rng.integers(0, 200)), int(rng.choice([4, 40, 1000, 70000, 2 ** 22, 2 ** 28]))
                strs.append(rle.string_from_counts(rng.integers(0, mag, m, dtype=np.int64).astype(np.uint32)))
            else:
                o = bytearray()
                for _ in range(int(rng.integers(0, 90))):
                    k = int(rng.integers(1, 11))
                    o += bytes((48 + (int(rng.integers(0, 32)) | 0x20)) for _ in range(k - 1)) + bytes([48 + int(rng.integers(0, 32))])
                strs.append(bytes(o))
        blob = b'x' * int(rng.integers(0, 4)) + b''.join(strs)
        off = len(blob) - sum(len(x) for x in strs)
        for x in strs:
            want = rle.counts_from_string(x) if len(x) else np.zeros(0, np.uint32)
            assert np.array_equal(decode(blob, off, len(x)), want), (t, len(x))
            off += len(x)


def test_model_of_the_flat_decode_pass_equals_the_oracle():
    """Lane-by-lane model (numpy) of the measure pass of rle_flat_crop_kernel (csrc/rle_flat.cu): the masks of a warp laid
    end to end as (0-run, 1-run) pairs, 32 pairs per pass; run ends and 1-pixel counts by segmented inclusive scans
    (a lane reaches min(lane, pair index in its mask) lanes down, the carry of lane 31 continues a mask into the next
    pass); first / last 1-pixel from two ballots; row ranges as 16-bit halves (min row | 0xffff - max row) under a
    per-half minimum; the segment's last lane adds to its mask; the painter's record per pair (column, first row,
    last row), or start | flag and length for runs that cross a column boundary.  Against area and tight box of the
    oracle's decoded mask, and the records against the mask itself."""
    from oracle import cocomask as rle
    rng = np.random.default_rng(9)
    SAT, SPAN = 0x7fffffff, 0x80000000

    def model(cnts, H, W):
        """cnts: list of uint32 run-count arrays (masks of one warp round) -> per-mask stats, per-pair records"""
        npairs = [(len(c) + 1) >> 1 for c in cnts]
        pb = np.concatenate([[0], np.cumsum(npairs)]).astype(np.int64)
        T, K = int(pb[-1]), len(cnts)
        area, first, last = np.zeros(K, np.int64), np.full(K, 0xffffffff, np.int64), np.zeros(K, np.int64)
        ymin, ymaxi, total = np.full(K, 0xffffffff, np.int64), np.full(K, 0xffffffff, np.int64), np.zeros(K, np.int64)
        rec, carry = [None] * T, 0
        for f0 in range(0, T, 32):
            lanes = np.arange(32)
            f = f0 + lanes
            ok = f < T
            j = np.searchsorted(pb, np.minimum(f, T - 1), side='right') - 1
            p = np.where(ok, f - pb[j], 0)
            z, o = np.zeros(32, np.int64), np.zeros(32, np.int64)
            for l in lanes[ok]:
                c = cnts[j[l]]
                z[l] = min(int(c[2 * p[l]]), SAT)
                o[l] = min(int(c[2 * p[l] + 1]), SAT) if 2 * p[l] + 1 < len(c) else 0
            end, osum = np.minimum(z + o, SAT), o.copy()
            reach = np.where(ok, np.minimum(lanes, p), 0)
            d = 1
            while d < 32:
                te, to = np.concatenate([np.zeros(d, np.int64), end[:-d]]), np.concatenate([np.zeros(d, np.int64), osum[:-d]])
                sel = d <= reach
                end, osum = np.where(sel, np.minimum(end + te, SAT), end), np.where(sel, osum + to, osum)
                d <<= 1
            end = np.where(ok & (p > lanes), np.minimum(end + carry, SAT), end)
            carry = int(end[31])
            inb = ok & (end <= H * W)
            good = inb & (o > 0)
            start = end - o
            ypk = np.full(32, 0xffffffff, np.int64)
            for l in lanes[ok]:
                if not good[l]:
                    rec[f[l]] = (0, 0x0000ffff)
                    continue
                x = int(start[l]) // H
                ys = int(start[l]) - x * H
                yl = ys + int(o[l]) - 1
                if yl >= H:
                    rec[f[l]] = (int(start[l]) | SPAN, int(o[l])); ypk[l] = (0xffff - (H - 1)) << 16
                else:
                    rec[f[l]] = (x, ys | (yl << 16)); ypk[l] = ys | ((0xffff - yl) << 16)
            d = 1
            while d < 32:
                t = np.concatenate([np.full(d, 0xffffffff, np.int64), ypk[:-d]])
                lo, hi = np.minimum(ypk & 0xffff, t & 0xffff), np.minimum(ypk >> 16, t >> 16)
                ypk = np.where(d <= reach, lo | (hi << 16), ypk)
                d <<= 1
            gm = sum(1 << int(l) for l in lanes[good])
            im = sum(1 << int(l) for l in lanes[inb])
            for l in lanes[ok]:
                seg_end = f[l] + 1 == pb[j[l] + 1]
                if not (seg_end or l == 31):
                    continue
                seg = ((2 << int(l)) - 1) & (0xffffffff << int(l - reach[l])) & 0xffffffff
                gs, i_s = gm & seg, im & seg
                if gs:
                    fl, ll, li = (gs & -gs).bit_length() - 1, gs.bit_length() - 1, i_s.bit_length() - 1
                    area[j[l]] += osum[li]
                    first[j[l]], last[j[l]] = min(first[j[l]], start[fl]), max(last[j[l]], end[ll])
                    ymin[j[l]], ymaxi[j[l]] = min(ymin[j[l]], ypk[l] & 0xffff), min(ymaxi[j[l]], ypk[l] >> 16)
                if seg_end:
                    total[j[l]] = end[l]
        return area, first, last, ymin, 0xffff - ymaxi, total, rec, pb

    done = 0
    for trial in range(40):
        H, W = int(rng.integers(3, 90)), int(rng.integers(3, 90))
        K = int(rng.integers(1, 7))
        masks = np.zeros((K, H, W), bool)
        for k in range(K):
            kind = rng.integers(0, 5)
            if kind == 0:
                continue                                                    # empty mask
            if kind == 1:
                masks[k] = rng.random((H, W)) < rng.random()              # noise: many runs, runs across columns
            elif kind == 2:
                masks[k, :, int(rng.integers(0, W)):] = True                # whole columns: one long run
            else:
                y, x, r = rng.integers(0, H), rng.integers(0, W), rng.integers(1, 12)
                yy, xx = np.ogrid[:H, :W]
                masks[k] = (yy - y) ** 2 + (xx - x) ** 2 <= r * r
        cnts = [rle.counts_from_string(rle.encode(np.asfortranarray(m.astype(np.uint8)))['counts']).astype(np.int64) for m in masks]
        if sum((len(c) + 1) >> 1 for c in cnts) > 256:                      # the kernel would split the round
            continue
        done += 1
        area, first, last, ymin, ymax, total, rec, pb = model(cnts, H, W)
        for k in range(K):
            m = masks[k]
            assert area[k] == m.sum() and total[k] == H * W, (trial, k)
            painted = np.zeros((W, H), bool)                                # column-major frame
            for fidx in range(int(pb[k]), int(pb[k + 1])):
                x, y = rec[fidx]
                if x & SPAN:
                    painted.reshape(-1)[(x & ~SPAN):(x & ~SPAN) + y] = True
                elif (y & 0xffff) <= (y >> 16):
                    painted[x, (y & 0xffff):(y >> 16) + 1] = True
            assert np.array_equal(painted.T, m), (trial, k)
            if m.any():
                ys_, xs_ = np.nonzero(m)
                assert first[k] // H == xs_.min() and (last[k] - 1) // H == xs_.max(), (trial, k)
                assert (ymin[k], ymax[k]) == (ys_.min(), ys_.max()), (trial, k)
    assert done >= 12


def test_grid_candidate_rule_sees_every_overlapping_pair_once():
    """Model (plain numpy) of the candidate search of csrc/intersect_grid.cu: columns binned into 32 x 32 clamped
    cells of side 2^shift, a row walks the cells of its box and takes a column only in the cell that holds the
    top-left corner of the two boxes' overlap.  Property: every pair of overlapping boxes is taken exactly once,
    no other pair ever -- for any shift, including boxes beyond the 32-cell extent (clamped) and huge boxes."""
    rng = np.random.default_rng(7)
    cell = lambda v, s: np.minimum(v >> s, 31)
    for trial in range(60):
        n_r, n_c = int(rng.integers(1, 40)), int(rng.integers(1, 60))
        size = int(rng.choice([40, 300, 2048, 5000]))
        def boxes(n):
            x0, y0 = rng.integers(0, size, n), rng.integers(0, size, n)
            big = rng.random(n) < 0.1
            wd = np.where(big, rng.integers(1, size, n), rng.integers(1, max(size // 20, 2), n))
            ht = np.where(big, rng.integers(1, size, n), rng.integers(1, max(size // 20, 2), n))
            return np.stack([x0, y0, np.minimum(x0 + wd, size) , np.minimum(y0 + ht, size)], 1)      # x1, y1 inclusive
        R, Cb = boxes(n_r), boxes(n_c)
        ext = int(max(Cb[:, 2].max(), Cb[:, 3].max()))
        mean = int(np.ceil((np.maximum(Cb[:, 2] - Cb[:, 0], Cb[:, 3] - Cb[:, 1]) + 1).mean()))
        shift = max(max(0, ext.bit_length() - 5), (max(mean, 1) - 1).bit_length())
        shift = int(rng.choice([shift, shift + 1, max(shift - 2, 0)]))          # the rule holds for any shift
        cells = {}
        for k, b in enumerate(Cb):
            for cy in range(cell(b[1], shift), cell(b[3], shift) + 1):
                for cx in range(cell(b[0], shift), cell(b[2], shift) + 1):
                    cells.setdefault((cx, cy), []).append(k)
        for r, a in enumerate(R):
            taken = []
            rcx0, rcy0 = cell(a[0], shift), cell(a[1], shift)
            for cy in range(rcy0, cell(a[3], shift) + 1):
                for cx in range(rcx0, cell(a[2], shift) + 1):
                    for k in cells.get((cx, cy), []):
                        b = Cb[k]
                        if b[0] <= a[2] and b[2] >= a[0] and b[1] <= a[3] and b[3] >= a[1] and \
                                max(rcx0, cell(b[0], shift)) == cx and max(rcy0, cell(b[1], shift)) == cy:
                            taken.append(k)
            want = [k for k, b in enumerate(Cb) if b[0] <= a[2] and b[2] >= a[0] and b[1] <= a[3] and b[3] >= a[1]]
            assert sorted(taken) == want, (trial, r, shift)


def test_match_bookkeeping_from_rows():
    from ampis_b200.analyze import _match_from_rows
    r = _match_from_rows(np.array([2, -1, 0, 2]), np.array([0.9, 0.0, 0.5, 0.6]), 4, 0.5)
    assert r['tp'].tolist() == [[0, 2], [3, 2]] and r['fn'].tolist() == [1, 2] and r['fp'].tolist() == [0, 1, 3]
    assert r['iou'].tolist() == [0.9, 0.6]
    r = _match_from_rows(np.array([-1]), np.array([0.0]), 2, 0.5)
    assert r['tp'].shape == (0,) and r['fp'].tolist() == [0, 1]


@pytest.mark.parametrize('name', ['c1_powder_example', 'c2_powder_batch', 'c3_satellites', 'c4_spheroidite'])
def test_synthetic_generator_is_valid_and_deterministic(name):
    from ampis_b200 import batch
    h1 = batch.synth(name, 2, 99, n_threads=1)
    h2 = batch.synth(name, 2, 99, n_threads=4)
    assert np.array_equal(h1.cnt, h2.cnt) and np.array_equal(h1.cnt_off, h2.cnt_off)
    hw = h1.h * h1.w
    for k in range(h1.n_masks):
        c = h1.cnt[h1.cnt_off[k]:h1.cnt_off[k] + h1.cnt_len[k]]
        assert int(c.astype(np.int64).sum()) == hw
        assert (c[1:] > 0).all()                         # canonical: no interior zero-length runs
    assert h1.n_masks == 2 * (h1.n_rows + h1.n_cols)


def test_get_ddicts_via2_json(tmp_path, monkeypatch):
    """data_utils.get_ddicts('via2') is host logic: VIA 2 JSON -> polygon data dicts (data_utils.py:436-480)."""
    import json
    from ampis_b200 import data_utils as D
    from ampis_b200.containers import BoxMode
    monkeypatch.chdir(tmp_path)
    (tmp_path / 'via').mkdir()
    via = {'_via_settings': {'core': {'default_filepath': '../images/'}},
           '_via_img_metadata': {
               'a.png123': {'filename': 'a.png', 'file_attributes': {'Size (width, height)': '1536, 1024', 'HFW': '103.6 um'},
                            'regions': [{'shape_attributes': {'name': 'polygon', 'all_points_x': [10, 20, 15],
                                                              'all_points_y': [5, 5, 30]}, 'region_attributes': {}},
                                        {'shape_attributes': {'name': 'polygon', 'all_points_x': [0, 4, 4, 0],
                                                              'all_points_y': [0, 0, 4, 4]}, 'region_attributes': {}}]},
               'b.png77': {'filename': 'b.png', 'file_attributes': {'Size (width, height)': '64, 32'}, 'regions': []}}}
    json.dump(via, open('via/anns.json', 'w'))
    dd = D.get_ddicts('via2', 'via/anns.json', dataset_class='Validation')
    assert [d['image_id'] for d in dd] == [0, 1] and dd[0]['file_name'] == 'via/../images/a.png'
    assert (dd[0]['height'], dd[0]['width'], dd[0]['HFW'], dd[0]['mask_format']) == (1024, 1536, '103.6 um', 'polygon')
    assert dd[0]['annotation_file'] == 'anns.json' and dd[0]['dataset_class'] == 'Validation'
    a = dd[0]['annotations']
    assert dd[0]['num_instances'] == 2 and a[0]['segmentation'] == [[10.5, 5.5, 20.5, 5.5, 15.5, 30.5]]
    assert a[0]['bbox'].tolist() == [10, 5, 20, 30] and a[0]['bbox_mode'] == BoxMode.XYXY_ABS and a[0]['category_id'] == 0
    assert dd[1]['num_instances'] == 0 and dd[1]['HFW'] is None and (dd[1]['height'], dd[1]['width']) == (32, 64)
    with pytest.raises(ValueError):
        D.get_ddicts('coco', 'via/anns.json')


@pytest.mark.timeout(600)
def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference runs without a GPU (the oracle port on the host cores) and prints the
    contract's JSON line; a non-zero rank exits silently."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0',
           '--config', 'c1_powder_example', '--cpu-images', '1']
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', RANK='0', WORLD_SIZE='1')
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=500)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better',
                'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['unit'] == 'pairs/s' and d['value'] > 0 and d['vs_baseline'] is None
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'model' not in d['config']
    r = subprocess.run(cmd, env=dict(env, RANK='1', WORLD_SIZE='2'), capture_output=True, text=True, timeout=100)
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_prediction_pickles_only_resolve_allowlisted_names():
    """containers.load_pickle: numpy arrays / scalars, plain containers and detectron2's Instances load; any other
    global -- builtins.eval, os.system, a numpy function -- is refused (the old filter let every builtin
    and every numpy.* name through)."""
    import io
    import os
    import pickle
    from ampis_b200.containers import Instances, _Unpickler
    x = {'a': np.arange(5, dtype=np.float32), 'b': np.float64(3.0), 'c': [np.int64(3)], 'd': (1, 2), 's': {1, 2},
         'i': Instances((4, 4), scores=np.ones(2))}
    y = _Unpickler(io.BytesIO(pickle.dumps(x))).load()
    assert np.array_equal(y['a'], x['a']) and y['b'] == 3.0 and y['s'] == {1, 2} and len(y['i']) == 2

    class Evil(object):
        def __init__(self, fn, arg):
            self.fn, self.arg = fn, arg

        def __reduce__(self):
            return (self.fn, (self.arg,))
    for fn, arg in ((eval, '1+1'), (os.system, 'true'), (getattr, 'x'), (np.load, 'nothing.npy'), (__import__, 'os')):
        with pytest.raises(pickle.UnpicklingError, match='refusing to load'):
            _Unpickler(io.BytesIO(pickle.dumps(Evil(fn, arg)))).load()


def _rows_from_oracle(gts, prs, mode):
    """An engine.ImagesRows filled by the ORACLE (what ampis_eval_images_host returns, computed on the CPU), so that
    the host-side bookkeeping over it can be tested without a GPU."""
    from ampis_b200 import engine
    from oracle import cocomask as rle
    r = engine.ImagesRows()
    n_img = len(gts)
    r.n_rows = np.array([len(g) for g in gts], np.int32)
    r.n_cols = np.array([len(p) for p in prs], np.int32)
    r.row_off = np.zeros(n_img + 1, np.int64)
    np.cumsum(r.n_rows, out=r.row_off[1:])
    r.mask_off = np.zeros(n_img + 1, np.int64)
    np.cumsum(r.n_rows.astype(np.int64) + r.n_cols, out=r.mask_off[1:])
    best_col, best_inter, best_score, area = [], [], [], []
    for g, p in zip(gts, prs):
        a_g, a_p = rle.area(g).astype(np.int64), rle.area(p).astype(np.int64)
        inter = np.array([[int(rle.merge_area(x, y, intersect=True)) for y in p] for x in g], np.int64).reshape(len(g), len(p))
        with np.errstate(invalid='ignore', divide='ignore'):
            score = np.where(inter > 0, inter / (a_g[:, None] + a_p[None, :] - inter), 0.0) if mode == 0 else \
                inter / a_g[:, None].astype(np.float64)
        if mode == 0:
            bc = np.where(score.max(axis=1) > 0, score.argmax(axis=1), -1)
            bs = score.max(axis=1)
        else:
            bc = inter.argmax(axis=1)               # first maximum of the intersections, as the kernel
            with np.errstate(invalid='ignore', divide='ignore'):
                bs = inter[np.arange(len(g)), bc] / a_g.astype(np.float64)
        best_col.append(bc)
        best_score.append(bs)
        best_inter.append(np.where(bc >= 0, inter[np.arange(len(g)), np.maximum(bc, 0)], 0))
        area += [a_g, a_p]
    r.best_col = np.concatenate(best_col).astype(np.int32)
    r.best_score = np.concatenate(best_score).astype(np.float64)
    r.best_inter = np.concatenate(best_inter).astype(np.uint32)
    r.area = np.concatenate(area).astype(np.uint32)
    return r


def test_batch_bookkeeping_equals_the_reference_loops():
    """The vectorised host bookkeeping of round 2 -- analyze._scores_from_rows_batch (det_seg_scores_batch),
    distributed._counts_from_rows (evaluate_sharded) and distributed._satellite_counts_from_rows
    (satellite_measurements_sharded) -- against the oracle's literal loops, image by image, with the per-row
    results supplied by the oracle instead of the GPU: ties, empty masks, images of different sizes."""
    from ampis_b200 import analyze, distributed
    from oracle import ampis_ref as R
    from oracle import cocomask as rle
    from tests import _util as U
    rng = np.random.default_rng(11)
    gts, prs = [], []
    for k in range(5):
        h, w = int(rng.integers(20, 60)), int(rng.integers(20, 60))
        m = U.rand_masks(rng, 9 + k + 7, h, w, p_empty=0.15)
        m[2] = m[9 + k]                                       # exact match
        m[9 + k + 1] = m[9 + k + 2] = m[3]                    # two identical predictions: the first one wins
        enc = [rle.encode(np.asfortranarray(x.astype(np.uint8))) for x in m]
        gts.append(enc[:9 + k])
        prs.append(enc[9 + k:])
    r = _rows_from_oracle(gts, prs, 0)
    for th in (0.0, 0.5, 0.8):
        got = analyze._scores_from_rows_batch(r, th)
        for g, p, res in zip(gts, prs, got):
            want = R.det_seg_scores(g, p, th)
            assert list(res) == list(want)
            for key in want:
                assert np.array_equal(np.asarray(res[key]), np.asarray(want[key]), equal_nan=True), (th, key)
                assert np.asarray(res[key]).dtype == np.asarray(want[key]).dtype, key
    ths = [0.3, 0.5, 0.75]
    counts = distributed._counts_from_rows(r, ths)
    for i, (g, p) in enumerate(zip(gts, prs)):
        for t, th in enumerate(ths):
            m = R.piecewise_rle_match(g, p, th)
            assert counts[i, t].tolist() == [len(m['tp']), len(m['fp']), len(m['fn'])]
    rs = _rows_from_oracle(gts, prs, 1)                       # "satellites" = gts, "particles" = prs
    sc, per = distributed._satellite_counts_from_rows(rs, 0.3)
    for i, (g, p) in enumerate(zip(gts, prs)):
        try:
            m = R.rle_satellite_match(p, g, 0.3)
            nm, pairs = len(m['satellite_matches']), m['match_pairs']
        except IndexError:
            nm, pairs = 0, {}
        assert sc[i].tolist() == [nm, len(g) - nm, len(pairs), len(p)]
        assert per[i].tolist() == [len(pairs[k]) for k in sorted(pairs)]
