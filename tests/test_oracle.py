"""CPU tests that pin the oracle (oracle/maskapi_ref.c, oracle/cocomask.py,
oracle/ampis_ref.py) -- SURVEY.md section 8c.  No GPU needed."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import ampis_ref as R
from oracle import cocomask as rle
from tests import _util as U


def _enc(a):
    return rle.encode(np.asfortranarray(np.array(a, np.uint8)))


def test_reference_known_answer():
    """The reference's only result-pinning test, analyze.py:702-728, verbatim values."""
    m1 = _enc([[1, 1, 0, 0], [1, 1, 0, 0], [0, 0, 0, 0], [0, 0, 0, 0]])
    m2 = _enc([[0, 0, 1, 1], [0, 0, 1, 1], [0, 0, 0, 0], [0, 0, 0, 0]])
    m3 = _enc([[0, 0, 0, 0], [0, 0, 0, 0], [1, 1, 0, 0], [1, 1, 0, 0]])
    m4 = _enc([[0, 0, 0, 0], [0, 0, 0, 0], [0, 0, 1, 1], [0, 0, 1, 1]])
    gt = [m1, m2, m3, m4]
    pred = [m3, m2, m4]
    assert np.all(R.piecewise_iou(gt, pred) == np.array([[0, 0, 0], [0, 1, 0], [1, 0, 0], [0, 0, 1]]))
    match = R.piecewise_rle_match(gt, pred)
    assert np.all(match['tp'] == np.array([[1, 1], [2, 0], [3, 2]]))
    assert np.all(match['fn'] == np.array([0]))
    assert np.all(match['fp'] == np.array([]))
    assert np.all(match['iou'] == np.ones(3))
    # m1 encodes to counts [0,2,2,2,10]
    assert list(rle.counts_from_string(m1['counts'])) == [0, 2, 2, 2, 10]


def _all_golden_strings():
    g = U.load('powder_match.npz')
    for k in range(len(g['names'])):
        _, gt, pr = U.powder_match_image(k)
        yield from gt
        yield from pr
    s = U.load('powder_satellite.npz')
    for k in range(len(s['names'])):
        yield from U.powder_satellite_image(k)[2]
    m = U.load('spheroidite_measure.npz')
    for k in range(len(m['names'])):
        yield from U.unpack_strings(m['%d_blob' % k], m['%d_off' % k], m['%d_size' % k])


def test_codec_on_real_pycocotools_strings():
    """The fixture strings were written by the real pycocotools.encode
    (data_utils.py:275): they are golden vectors for rleFrString / rleToString."""
    n = 0
    for m in _all_golden_strings():
        c = rle.counts_from_string(m['counts'])
        h, w = m['size']
        assert int(c.astype(np.int64).sum()) == h * w
        assert (c[1:-1] > 0).all()            # canonical encode output has no interior zero runs
        assert rle.string_from_counts(c) == m['counts']
        n += 1
    assert n > 3000


def test_encode_canonical_form_matches_real_strings():
    """decode -> encode must give back the exact bytes pycocotools wrote."""
    _, gt, _ = U.powder_match_image(0)
    for m in gt[:40]:
        d = rle.decode(m)
        assert rle.encode(np.asfortranarray(d))['counts'] == m['counts']
        assert int(d.sum()) == int(rle.area(m))


def test_runwalk_iou_equals_dense_formulation():
    """Second, independent formulation: packed AND + popcount, IoU = I/(a+b-I)."""
    g, gt, pr = U.powder_match_image(0)
    hw = int(g['0_size'][0]) * int(g['0_size'][1])
    pg = [U.counts_to_packed(rle.counts_from_string(m['counts']), hw) for m in gt]
    pp = [U.counts_to_packed(rle.counts_from_string(m['counts']), hw) for m in pr]
    ag = np.array([U.popcount(x) for x in pg])
    ap = np.array([U.popcount(x) for x in pp])
    assert (ag == g['0_gt_area']).all() and (ap == g['0_pr_area']).all()
    iou = R.piecewise_iou(gt, pr)
    nz = g['0_iou_nz_idx']
    assert (np.argwhere(iou > 0) == nz).all()
    rng = np.random.default_rng(0)
    extra = np.stack([rng.integers(0, len(gt), 300), rng.integers(0, len(pr), 300)], 1)
    for i, j in np.concatenate([nz, extra]):
        inter = U.popcount(pg[i] & pp[j])
        want = inter / (ag[i] + ap[j] - inter) if inter else 0.0
        assert iou[i, j] == want
        assert int(rle.merge_area(gt[i], pr[j])) == inter
        assert int(rle.area(rle.merge([gt[i], pr[j]], intersect=True))) == inter
        assert int(rle.area(rle.merge([gt[i], pr[j]], intersect=False))) == ag[i] + ap[j] - inter


def test_golden_regression_match_and_scores():
    for k in (0, 3):
        g, gt, pr = U.powder_match_image(k)
        res = R.det_seg_scores(gt, pr, 0.5)
        for key in ('det_tp', 'det_fn', 'det_fp', 'seg_tp', 'seg_fn', 'seg_fp', 'det_tp_iou',
                    'seg_precision', 'seg_recall'):
            assert np.array_equal(np.asarray(res[key]), g['%d_%s' % (k, key)]), key
        assert [res['det_precision'], res['det_recall']] == list(g['%d_det_pr' % k])


def test_golden_regression_satellites():
    s, part, sat = U.powder_satellite_image(1)
    res = R.rle_satellite_match(part, sat, 0.5)
    for key in ('satellite_matches', 'satellites_unmatched', 'particles_unmatched', 'intersection_scores'):
        assert np.array_equal(res[key], s['1_%s' % key])
    # satellite score = popcount(AND)/area(sat)
    hw = int(s['1_size'][0]) * int(s['1_size'][1])
    for (si, pi), sc in list(zip(res['satellite_matches'], res['intersection_scores']))[:25]:
        a = U.counts_to_packed(rle.counts_from_string(sat[si]['counts']), hw)
        b = U.counts_to_packed(rle.counts_from_string(part[pi]['counts']), hw)
        assert sc == U.popcount(a & b) / U.popcount(a)


def test_polygon_rectangles_closed_form():
    """Axis-aligned rectangles with half-integer corners cover exactly the pixels whose
    centres lie inside (SURVEY A.7 sanity case), any vertex order / start."""
    r = rle.frPyObjects([[2.5, 1.5, 6.5, 1.5, 6.5, 4.5, 2.5, 4.5]], 8, 10)[0]
    assert list(rle.counts_from_string(r['counts'])) == [26, 3, 5, 3, 5, 3, 5, 3, 27]
    rng = np.random.default_rng(1)
    for _ in range(50):
        h, w = rng.integers(4, 40, 2)
        x0, x1 = sorted(rng.integers(0, w, 2)); y0, y1 = sorted(rng.integers(0, h, 2))
        x1 += 1; y1 += 1
        pts = [(x0 - .5, y0 - .5), (x1 - .5, y0 - .5), (x1 - .5, y1 - .5), (x0 - .5, y1 - .5)]
        sh = rng.integers(0, 4)
        pts = pts[sh:] + pts[:sh]
        if rng.random() < .5:
            pts = pts[::-1]
        poly = [c + 0.5 for p in pts for c in p]          # pixel-edge coordinates x0..x1, y0..y1
        m = rle.decode(rle.frPyObjects([poly], int(h), int(w))[0]).astype(bool)
        want = np.zeros((h, w), bool)
        want[y0:y1, x0:x1] = True
        assert (m == want).all()


def test_polygon_golden_vs_shoelace():
    p = U.load('powder_polygons.npz')
    xy, off, area = p['0_poly_xy'], p['0_poly_off'], p['0_area']
    size = tuple(int(v) for v in p['0_size'])
    rl = U.unpack_strings(p['0_rle_blob'], p['0_rle_off'], size)
    rel = []
    for i in range(len(off) - 1):
        q = xy[off[i]:off[i + 1]]
        sh = R.shoelace_area(q[::2], q[1::2])
        rel.append((area[i] - sh) / sh)
        if i < 25:   # regression of the frozen strings
            assert rle.frPyObjects([q], *size)[0]['counts'] == rl[i]['counts']
    rel = np.array(rel)
    assert abs(np.median(rel)) < 0.005 and np.percentile(np.abs(rel), 90) < 0.05


def test_measurements_regression():
    m = U.load('spheroidite_measure.npz')
    masks = U.unpack_strings(m['0_blob'], m['0_off'], m['0_size'])
    bm = R.rle_to_bitmask_array(masks)
    assert (bm.sum(axis=(1, 2)) == m['0_area']).all()
    assert np.array_equal(R.extract_boxes(bm), m['0_boxes_d2'])
    assert np.array_equal(R.extract_boxes(bm, box_mode='matterport'), m['0_boxes_mp'])
    assert np.array_equal(R.edge_inliers(masks, tuple(m['0_size']), 1), m['0_edge_inliers'])
    # rleToBbox contains the tight box
    bb, t = m['0_rlebbox'], m['0_boxes_d2']
    ne = m['0_area'] > 0
    assert (bb[ne, 0] <= t[ne, 0]).all() and (bb[ne, 0] + bb[ne, 2] - 1 >= t[ne, 2]).all()
    assert (bb[ne, 1] <= t[ne, 1]).all() and (bb[ne, 1] + bb[ne, 3] - 1 >= t[ne, 3]).all()


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 17), st.integers(1, 19), st.integers(0, 2 ** 31 - 1))
def test_properties_random_masks(h, w, seed):
    rng = np.random.default_rng(seed)
    a = rng.random((h, w)) < rng.random()
    b = rng.random((h, w)) < rng.random()
    ra, rb = _enc(a), _enc(b)
    assert (rle.decode(ra).astype(bool) == a).all()                       # decode(encode) = id
    assert int(rle.area(ra)) == int(a.sum())                              # area = popcount
    assert int(rle.area(rle.merge([ra, rb], intersect=True))) == int((a & b).sum())
    assert int(rle.area(rle.merge([ra, rb], intersect=False))) == int((a | b).sum())
    i, u = int((a & b).sum()), int((a | b).sum())
    assert rle.iou([ra], [rb], [False])[0, 0] == (i / u if i else 0.0)
    c = rle.counts_from_string(ra['counts'])
    assert rle.string_from_counts(c) == ra['counts']
    if a.any():
        bb = rle.to_bbox(ra)
        ys, xs = np.where(a)
        assert bb[0] <= xs.min() and bb[0] + bb[2] - 1 >= xs.max()
        assert bb[1] <= ys.min() and bb[1] + bb[3] - 1 >= ys.max()


def test_match_edge_cases():
    z = _enc(np.zeros((6, 5)))
    o = _enc(np.ones((6, 5)))
    half = _enc(np.pad(np.ones((3, 5)), ((0, 3), (0, 0))))
    # zero-area masks give IoU exactly 0.0, never NaN
    assert (R.piecewise_iou([z, o], [z, half]) == np.array([[0, 0], [0, 0.5]])).all()
    m = R.piecewise_rle_match([z, o, half], [half, half, z], 0.4)
    assert m['tp'].tolist() == [[1, 0], [2, 0]] and m['fn'].tolist() == [0] and m['fp'].tolist() == [1, 2]
    # strict '>' on the threshold
    assert len(R.piecewise_rle_match([o], [half], 0.5)['tp']) == 0
    # nothing matches: tp has shape (0,)
    assert R.piecewise_rle_match([z], [z])['tp'].shape == (0,)
    with pytest.raises(ZeroDivisionError):
        R.det_seg_scores([], [], 0.5)
    # satellites: zero-area satellite -> NaN row -> unmatched; no match at all -> IndexError
    res = R.rle_satellite_match([half, o], [z, half], 0.5)
    assert res['satellite_matches'].tolist() == [[1, 0]] and res['satellites_unmatched'].tolist() == [0]
    assert res['particles_unmatched'].tolist() == [1] and res['match_pairs'] == {0: [1]}
    with pytest.raises(IndexError):
        R.rle_satellite_match([half], [z], 0.5)


def test_regionprops_restatement_closed_forms():
    """Known answers for the skimage restatement (oracle/ampis_ref.regionprops_one): filled rectangles
    have closed-form moments, skimage's 4-neighbourhood perimeter of an a x b rectangle is 2(a-1)+2(b-1),
    a one-pixel-wide line of n pixels has perimeter n-2... (skimage documents perimeter(10x10 square) = 36)."""
    from oracle import ampis_ref as R
    m = np.zeros((40, 50), bool)
    m[5:15, 10:20] = True                                         # 10 x 10 square
    p = R.regionprops_one(m)
    assert p['area'] == 100 and p['bbox'] == (5, 10, 15, 20) and p['bbox_area'] == 100
    assert p['perimeter'] == 36.0 and p['convex_area'] == 100 and p['solidity'] == 1.0 and p['extent'] == 1.0
    assert p['centroid'] == (9.5, 14.5) and p['local_centroid'] == (4.5, 4.5)
    assert np.isclose(p['major_axis_length'], 4 * np.sqrt(99 / 12)) and np.isclose(p['minor_axis_length'], 4 * np.sqrt(99 / 12))
    assert np.isclose(p['eccentricity'], 0.0, atol=1e-7) and np.isclose(p['equivalent_diameter'], np.sqrt(400 / np.pi))
    m[:] = False
    m[8:12, 3:33] = True                                          # 4 rows x 30 columns: long axis along the columns
    p = R.regionprops_one(m)
    assert p['perimeter'] == 2 * 3 + 2 * 29
    assert np.isclose(p['major_axis_length'], 4 * np.sqrt((30 ** 2 - 1) / 12))
    assert np.isclose(p['minor_axis_length'], 4 * np.sqrt((4 ** 2 - 1) / 12))
    assert np.isclose(abs(p['orientation']), np.pi / 2)           # skimage: angle to the row axis
    assert np.isclose(R.regionprops_one(m.T)['orientation'], 0.0)
    m[:] = False
    m[7, 2:12] = True                                             # a 10-pixel line
    p = R.regionprops_one(m)
    assert p['perimeter'] == 8.0 and p['convex_area'] == 10       # end pixels (one border neighbour) carry no weight
    assert p['minor_axis_length'] == 0.0 and p['eccentricity'] == 1.0
    m[:] = False
    m[[3, 4, 4], [3, 3, 4]] = True                                # three-pixel corner: its hull holds no fourth pixel centre
    assert R.regionprops_one(m)['convex_area'] == 3
    m[:] = False
    m[10:21, 10] = m[10, 10:21] = m[20, 10:21] = m[10:21, 20] = True      # hollow square ring: hull = filled square
    p = R.regionprops_one(m)
    assert p['area'] == 40 and p['convex_area'] == 121 and np.isclose(p['solidity'], 40 / 121)
    assert R.regionprops_one(np.zeros((5, 5), bool)) == {}


def test_polygon2mask_and_labelling_restatements_known_answers():
    from oracle import ampis_ref as R
    # an axis-aligned square with corners on pixel centres: the crossing rule keeps the top / left edges
    sq = [2.0, 3.0, 6.0, 3.0, 6.0, 7.0, 2.0, 7.0]                 # x0,y0,... : x in [2,6], y in [3,7]
    mask = R.poly2mask([sq], (10, 10))[0]
    assert mask.sum() == 16 and mask[3:7, 2:6].all()
    # 8-connectivity, raster numbering
    img = np.array([[1, 0, 0, 1],
                    [0, 1, 0, 1],
                    [0, 0, 0, 0],
                    [1, 1, 0, 1]], bool)
    lab = R.label_binary(img)
    assert lab.tolist() == [[1, 0, 0, 2], [0, 1, 0, 2], [0, 0, 0, 0], [3, 3, 0, 4]]
    anns = R.annotations_from_label_image(img, binary=True)
    assert [a[0].tolist() for a in anns] == [[0, 0, 1, 1], [3, 0, 3, 1], [0, 3, 1, 3], [3, 3, 3, 3]]
    lab2 = np.array([[0, 7, 7], [3, 3, 0], [0, 0, 9]])
    anns = R.annotations_from_label_image(lab2, binary=False)
    assert [a[0].tolist() for a in anns] == [[0, 1, 1, 1], [1, 0, 2, 0], [2, 2, 2, 2]]       # values 3, 7, 9


def test_edge_distance_and_class_maps_known_answers():
    from oracle import ampis_ref as R
    from oracle import cocomask as rle
    g = np.zeros((8, 8), np.uint8)
    p = np.zeros((8, 8), np.uint8)
    g[2:6, 2:6] = 1
    p[2:6, 3:7] = 1                                                # shifted one pixel to the right
    ge, pe = [rle.encode(np.asfortranarray(g))], [rle.encode(np.asfortranarray(p))]
    fp, fn = R.mask_edge_distance(ge, pe, [np.array([2, 6, 2, 6])], [np.array([2, 6, 3, 7])], np.array([[0, 0]]))
    assert fp[0].tolist() == [1.0] * 4 and fn[0].tolist() == [1.0] * 4
    masks = R.seg_perf_masks(ge, pe, np.array([[0, 0]]), 'reduced')
    assert [int(rle.area(m)) for m in masks] == [12, 4, 4, 0]     # TP, FN, FP, overlap classes
    assert R.merge_boxes([2, 6, 2, 6], [2, 6, 3, 7]).tolist() == [2, 6, 2, 7]


def test_one_shot_oracle_forms_equal_the_literal_loops():
    """tests/_util.py's one-shot forms (one rleIou call over all pairs + merges on overlapping pairs only), which the
    native-size GPU tests use as the checker, against the literal restatement of the reference's loops
    (oracle/ampis_ref.py) on the shipped powder fixtures and on random masks with empty, tied and full-frame cases."""
    for k in (0, 2):
        _, gt, pr = U.powder_match_image(k)
        iou = U.iou_matrix_one_shot(rle, gt, pr)
        assert np.array_equal(iou, R.piecewise_iou(gt, pr))
        for th in (0.5, 0.75, 0.95):
            want, got = R.det_seg_scores(gt, pr, th), U.det_seg_scores_one_shot(rle, gt, pr, th, iou=iou)
            assert want.keys() == got.keys()
            for key in want:
                assert np.array_equal(np.asarray(want[key]), np.asarray(got[key])), (k, th, key)
    s, part, sat = U.powder_satellite_image(1)
    want, got = R.rle_satellite_match(part, sat, 0.5), U.satellite_match_one_shot(rle, part, sat, 0.5)
    for key in ('satellite_matches', 'satellites_unmatched', 'particles_unmatched', 'intersection_scores'):
        assert np.array_equal(want[key], got[key]), key
    assert {k_: list(v) for k_, v in want['match_pairs'].items()} == {k_: list(v) for k_, v in got['match_pairs'].items()}
    rng = np.random.default_rng(5)
    for trial in range(6):
        h, w = int(rng.integers(3, 60)), int(rng.integers(3, 60))
        m = U.rand_masks(rng, 22, h, w, p_empty=0.2)
        m[3] = m[4]                                    # exact ties: the first arg-max must win
        m[15] = m[16] = m[4]
        m[7] = True
        enc = [rle.encode(np.asfortranarray(x.astype(np.uint8))) for x in m]
        gt, pr = enc[:9], enc[9:]
        for th in (0.0, 0.3, 0.5):
            try:
                want = R.det_seg_scores(gt, pr, th)
            except ZeroDivisionError:
                with pytest.raises(ZeroDivisionError):
                    U.det_seg_scores_one_shot(rle, gt, pr, th)
                continue
            got = U.det_seg_scores_one_shot(rle, gt, pr, th)
            for key in want:
                assert np.array_equal(np.asarray(want[key]), np.asarray(got[key]), equal_nan=True), (trial, th, key)
        try:
            want = R.rle_satellite_match(pr, gt, 0.4)
        except IndexError:
            with pytest.raises(IndexError):
                U.satellite_match_one_shot(rle, pr, gt, 0.4)
            continue
        got = U.satellite_match_one_shot(rle, pr, gt, 0.4)
        for key in ('satellite_matches', 'satellites_unmatched', 'particles_unmatched', 'intersection_scores'):
            assert np.array_equal(want[key], got[key]), (trial, key)
