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
    q.put((rank, r['totals'], r['per_image'], r['index'], {k: v for k, v in s.items() if k != 'index'}, h))
    dist.barrier()
    dist.destroy_process_group()


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
    for rank, totals, per_image, index, sat, hist in res:
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
    q.put((rank, r['totals'], r['per_image'], r['index'], {k: v for k, v in s.items() if k != 'index'}, h))
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
    for rank, totals, per_image, index, sat, hist in sorted(res, key=lambda x: x[0]):
        assert np.array_equal(index, np.arange(rank, n_img, 2))
        assert np.array_equal(totals, want.sum(axis=0)) and np.array_equal(per_image, want)
        assert sat['n_satellites'] == want_sat[:, 0].sum() and sat['n_satellited_particles'] == want_sat[:, 2].sum()
        assert np.array_equal(sat['spp_hist'], want_hist)
        assert np.array_equal(hist, np.bincount(np.clip(areas // 50, 0, 7), minlength=8))
