"""World-size-2 tests of the sharded evaluation on CPU (gloo).  The per-shard computation is
supplied by the CPU oracle here (the GPU pipeline is covered by test_gpu_parity.py); what is
tested is the N>1 host path: strided sharding, the all-reduce payload, gathering in image order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dataset(n_img):
    from ampis_b200 import batch
    from oracle import cocomask as rle
    cfg = dict(batch.CONFIGS['c2_powder_batch'], h=96, w=80, n_rows=9, n_cols=8, median_diam=14.0)
    host = batch.synth(cfg, n_img, 321)
    gts, prs = [], []
    for g in range(n_img):
        rows, cols = host.image_masks(g)
        gts.append([{'size': [96, 80], 'counts': rle.string_from_counts(c)} for c in rows])
        prs.append([{'size': [96, 80], 'counts': rle.string_from_counts(c)} for c in cols])
    return gts, prs


def _oracle_counts(gt_shard, pr_shard, thresholds):
    from oracle import ampis_ref as R
    out = np.zeros((len(gt_shard), len(thresholds), 3), np.int64)
    for i, (g, p) in enumerate(zip(gt_shard, pr_shard)):
        for t, th in enumerate(thresholds):
            m = R.piecewise_rle_match(g, p, th)
            out[i, t] = [len(m['tp']), len(m['fp']), len(m['fn'])]
    return out


def _oracle_sat(part_shard, sat_shard, thresh, n_bins):
    from oracle import ampis_ref as R
    counts = np.zeros((len(part_shard), 4), np.int64)
    hist = np.zeros(n_bins, np.int64)
    for i, (p, s) in enumerate(zip(part_shard, sat_shard)):
        try:
            m = R.rle_satellite_match(p, s, thresh)
            nm, pairs = len(m['satellite_matches']), m['match_pairs']
        except IndexError:
            nm, pairs = 0, {}
        counts[i] = [nm, len(s) - nm, len(pairs), len(p)]
        for v in pairs.values():
            hist[min(len(v), n_bins - 1)] += 1
    return counts, hist


def _oracle_sat_rows(part_shard, sat_shard, thresh):
    """CPU stand-in of distributed._satellite_rows: counts and satellites per satellited particle from the oracle."""
    from oracle import ampis_ref as R
    counts = np.zeros((len(part_shard), 4), np.int64)
    per = []
    for i, (p, s) in enumerate(zip(part_shard, sat_shard)):
        try:
            m = R.rle_satellite_match(p, s, thresh)
            nm, pairs = len(m['satellite_matches']), m['match_pairs']
        except IndexError:
            nm, pairs = 0, {}
        counts[i] = [nm, len(s) - nm, len(pairs), len(p)]
        per.append(np.asarray([len(pairs[k]) for k in sorted(pairs)], np.int64))
    return counts, per


def _powder_images(gts, prs):
    """PowderSatelliteImage objects over the synthetic RLE lists (particles = prs, satellites = gts), with a
    different horizontal field width per image so that psd scales every image by its own factor."""
    from ampis_b200.applications.powder import PowderSatelliteImage
    from ampis_b200.containers import Instances
    from ampis_b200.structures import InstanceSet, RLEMasks
    out = []
    for k, (g, p) in enumerate(zip(gts, prs)):
        sets = []
        for masks in (p, g):
            iset = InstanceSet()
            iset.instances = Instances((96, 80), masks=RLEMasks(masks), class_idx=np.zeros(len(masks), int))
            iset.HFW, iset.HFW_units = 100.0 + 7 * k, 'um'
            sets.append(iset)
        out.append(PowderSatelliteImage(particles=sets[0], satellites=sets[1]))
    return out


def _oracle_areas(item):
    from oracle import cocomask as rle
    from ampis_b200.structures import masks_to_rle
    return rle.area(masks_to_rle(item))


def _sharded_measurements(D, gts, prs):
    psi = _powder_images(gts, prs)
    curves = {(x, y, d): D.psd_sharded(psi, xvals=x, yvals=y, distance=d, areas_fn=_oracle_areas)
              for x, y, d in (('d_eq', 'cvf', 'length'), ('area', 'counts', 'pixels'))}
    sm = D.satellite_measurements_sharded(psi, thresh=0.3, match_fn=_oracle_sat_rows)
    return curves, sm


def _worker(rank, world, port, n_img, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from ampis_b200 import distributed as D
    gts, prs = _dataset(n_img)
    th = [0.5, 0.75]
    r = D.evaluate_sharded(gts, prs, th, compute_fn=_oracle_counts)
    s = D.satellites_sharded(prs, gts, 0.3, 16, compute_fn=_oracle_sat)
    h = D.area_histogram_sharded(gts, 0, 50, 8,
                                 compute_fn=lambda shard: np.bincount(np.clip(np.concatenate(
                                     [__import__('oracle.cocomask', fromlist=['x']).area(m) for m in shard]
                                     or [np.zeros(0, np.int64)]).astype(np.int64) // 50, 0, 7), minlength=8))
    curves, sm = _sharded_measurements(D, gts, prs)
    q.put((rank, r['totals'], r['per_image'], r['index'], {k: v for k, v in s.items() if k != 'index'}, h, curves, sm))
    dist.barrier()
    dist.destroy_process_group()


def _check_sharded_measurements(gts, prs, curves, sm):
    """psd_sharded / satellite_measurements_sharded of any rank == the single-process dicts, bit for bit
    (single process: powder's own code fed by the oracle's areas / matches, and the oracle's restatement)."""
    from ampis_b200.applications import powder
    from oracle import ampis_ref as R
    from oracle import cocomask as rle
    psi = _powder_images(gts, prs)
    for (x, y, d), got in curves.items():
        want = powder.psd(psi, xvals=x, yvals=y, distance=d, plot=False, return_results=True,
                          _areas_of=lambda items: [_oracle_areas(i) for i in items])
        assert got.keys() == want.keys() and got['x_label'] == want['x_label'] and got['y_label'] == want['y_label']
        assert np.array_equal(got['x'], want['x']) and np.array_equal(got['y'], want['y'])
        scale = [(100.0 + 7 * k) / 80 for k in range(len(prs))]
        areas = [rle.area(p) * c ** 2 for p, c in zip(prs, scale)] if d == 'length' else [rle.area(p) for p in prs]
        ref = R.psd_from_areas(areas, x, y)
        assert np.array_equal(got['x'], ref['x']) and np.array_equal(got['y'], ref['y'])
    matches = []
    for p, s in zip(prs, gts):
        try:
            matches.append(R.rle_satellite_match(p, s, 0.3))
        except IndexError:
            matches.append({'match_pairs': {}, 'particles_unmatched': np.arange(len(p)),
                            'satellites_unmatched': np.arange(len(s))})
    want = R.satellite_measurements(matches, [len(p) for p in prs], [len(s) for s in gts])
    assert list(sm.keys()) == list(want.keys())
    for k in want:
        assert np.array_equal(np.asarray(sm[k]), np.asarray(want[k])), k


@pytest.mark.timeout(300)
def test_two_rank_gloo_matches_single_process():
    n_img = 7
    gts, prs = _dataset(n_img)
    th = [0.5, 0.75]
    want = _oracle_counts(gts, prs, th)
    want_sat, want_hist = _oracle_sat(prs, gts, 0.3, 16)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_img, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    res.sort(key=lambda x: x[0])
    for rank, totals, per_image, index, sat, hist, curves, sm in res:
        _check_sharded_measurements(gts, prs, curves, sm)
        assert np.array_equal(index, np.arange(rank, n_img, 2))
        assert np.array_equal(totals, want.sum(axis=0))
        assert np.array_equal(per_image, want)
        assert sat['n_images'] == n_img and sat['n_particles'] == sum(len(p) for p in prs)
        assert sat['n_satellites'] == want_sat[:, 0].sum() and sat['n_satellites_unmatched'] == want_sat[:, 1].sum()
        assert sat['n_satellited_particles'] == want_sat[:, 2].sum()
        assert np.array_equal(sat['spp_hist'], want_hist)
        assert hist.sum() == sum(len(g) for g in gts)
    assert np.array_equal(res[0][5], res[1][5])


def test_single_process_paths():
    from ampis_b200 import distributed as D
    assert D.world_info() == (0, 1)
    assert D.shard_indices(10, 1, 4).tolist() == [1, 5, 9]
    rows = D.gather_rows(np.arange(6).reshape(3, 2), np.array([0, 1, 2]), 3)
    assert rows.tolist() == [[0, 1], [2, 3], [4, 5]]
    gts, prs = _dataset(3)
    r = D.evaluate_sharded(gts, prs, [0.5], compute_fn=_oracle_counts)
    assert np.array_equal(r['per_image'], _oracle_counts(gts, prs, [0.5]))
    assert [a.tolist() for a in D.all_gather_varlen([3, 1, 2])] == [[3, 1, 2]]
    _check_sharded_measurements(gts, prs, *_sharded_measurements(D, gts, prs))


@pytest.mark.gpu
def test_sharded_gpu_pipeline_single_rank():
    """The default compute functions (GPU batch pipeline) against the oracle, world size 1."""
    from ampis_b200 import distributed as D
    gts, prs = _dataset(5)
    th = [0.5, 0.75, 0.9]
    r = D.evaluate_sharded(gts, prs, th)
    assert np.array_equal(r['per_image'], _oracle_counts(gts, prs, th))
    s = D.satellites_sharded(prs, gts, 0.3, 16)
    want_sat, want_hist = _oracle_sat(prs, gts, 0.3, 16)
    assert s['n_satellites'] == want_sat[:, 0].sum() and s['n_satellited_particles'] == want_sat[:, 2].sum()
    assert np.array_equal(s['spp_hist'], want_hist)
    from oracle import cocomask as rle
    areas = np.concatenate([rle.area(g) for g in gts]).astype(np.int64)
    assert np.array_equal(D.area_histogram_sharded(gts, 0, 50, 8), np.bincount(np.clip(areas // 50, 0, 7), minlength=8))
    # exact-value PSD and the satellite summary through the GPU measurement / matching of this rank
    psi = _powder_images(gts, prs)
    curves = {(x, y, d): D.psd_sharded(psi, xvals=x, yvals=y, distance=d)
              for x, y, d in (('d_eq', 'cvf', 'length'), ('area', 'counts', 'pixels'))}
    _check_sharded_measurements(gts, prs, curves, D.satellite_measurements_sharded(psi, thresh=0.3))
    from ampis_b200.applications import powder
    for im in psi:
        im.compute_matches(0.3)
    single = powder.satellite_measurements(psi, print_summary=False, output_dict=True)
    sharded = D.satellite_measurements_sharded(psi, thresh=0.3)
    assert all(np.array_equal(np.asarray(single[k]), np.asarray(sharded[k])) for k in single)


def _nccl_worker(rank, world, port, n_img, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from ampis_b200 import distributed as D
    gts, prs = _dataset(n_img)
    th = [0.5, 0.75]
    r = D.evaluate_sharded(gts, prs, th)                     # GPU pipeline on this rank's device + NCCL all-reduce
    s = D.satellites_sharded(prs, gts, 0.3, 16)
    h = D.area_histogram_sharded(gts, 0, 50, 8)
    psi = _powder_images(gts, prs)
    curves = {(x, y, d): D.psd_sharded(psi, xvals=x, yvals=y, distance=d)
              for x, y, d in (('d_eq', 'cvf', 'length'), ('area', 'counts', 'pixels'))}
    sm = D.satellite_measurements_sharded(psi, thresh=0.3)
    q.put((rank, r['totals'], r['per_image'], r['index'], {k: v for k, v in s.items() if k != 'index'}, h, curves, sm))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_two_rank_nccl_on_two_gpus():
    """Same check as the gloo test with the real thing: two processes, two GPUs, the CUDA pipeline on
    each shard and NCCL for the reduction / gather.  Skipped on a single-GPU box."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    n_img = 7
    gts, prs = _dataset(n_img)
    th = [0.5, 0.75]
    want = _oracle_counts(gts, prs, th)
    want_sat, want_hist = _oracle_sat(prs, gts, 0.3, 16)
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, n_img, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=500) for _ in procs]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    from oracle import cocomask as rle
    areas = np.concatenate([rle.area(g) for g in gts]).astype(np.int64)
    for rank, totals, per_image, index, sat, hist, curves, sm in sorted(res, key=lambda x: x[0]):
        _check_sharded_measurements(gts, prs, curves, sm)
        assert np.array_equal(index, np.arange(rank, n_img, 2))
        assert np.array_equal(totals, want.sum(axis=0)) and np.array_equal(per_image, want)
        assert sat['n_satellites'] == want_sat[:, 0].sum() and sat['n_satellited_particles'] == want_sat[:, 2].sum()
        assert np.array_equal(sat['spp_hist'], want_hist)
        assert np.array_equal(hist, np.bincount(np.clip(areas // 50, 0, 7), minlength=8))
