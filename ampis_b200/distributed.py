"""Multi-GPU evaluation: one process per GPU, images sharded across ranks, one small
all-reduce at the end.

Images are independent units of the hot path, so there is no data-path collective: rank r
evaluates images r, r+W, r+2W, ... on its own GPU.  The only exchange is the sum of the
dataset-level counters -- [TP, FP, FN] per IoU threshold, satellite totals, histograms -- an
int64 vector of a few hundred bytes, reduced with torch.distributed (NCCL over NVLink on the
B200 box, gloo in the CPU tests).  Per-image results stay rank-local unless gathered.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import batch, engine


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_indices(n_items, rank, world):
    """Strided shard: rank r owns items r, r+W, ... (order restored by index on gather)."""
    return np.arange(rank, n_items, world, dtype=np.int64)


def all_reduce_sum_(t, group=None):
    """In-place SUM all-reduce of an int64 tensor (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def gather_rows(local_rows, local_index, n_items, group=None):
    """Gather per-image rows (int64 [n_local, k]) of all ranks into [n_items, k] ordered by
    image index; every rank receives the result."""
    rank, world = world_info(group)
    local_rows = np.asarray(local_rows, np.int64).reshape(len(local_index), -1)
    k = local_rows.shape[1]
    out = np.zeros((n_items, k), np.int64)
    if world == 1:
        out[local_index] = local_rows
        return out
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else 'cpu'
    n_max = (n_items + world - 1) // world
    buf = torch.full((n_max, k + 1), -1, dtype=torch.int64, device=dev)
    if len(local_index):
        buf[:len(local_index), 0] = torch.from_numpy(np.asarray(local_index, np.int64)).to(dev)
        buf[:len(local_index), 1:] = torch.from_numpy(local_rows).to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf, group=group)
    for p in parts:
        p = p.cpu().numpy()
        ok = p[:, 0] >= 0
        out[p[ok, 0]] = p[ok, 1:]
    return out


def _gpu_match_counts(gt_lists, pred_lists, thresholds):
    """[n_local, T, 3] TP/FP/FN of the local images through the batch pipeline on this rank's GPU."""
    n = len(gt_lists)
    T = len(thresholds)
    if n == 0:
        return np.zeros((0, T, 3), np.int64)
    masks = []
    for g, p in zip(gt_lists, pred_lists):
        masks += list(g) + list(p)
    table = engine.table_from_rle(masks, layout=engine.LAYOUT_CROP)     # smallest storage, same results
    groups = engine.Groups.interleaved(table.device, [len(g) for g in gt_lists], [len(p) for p in pred_lists])
    rows = engine.intersect_rows(table, groups, engine.MODE_IOU)
    counts, _ = engine.match_counts(rows, groups, thresholds)
    return counts.cpu().numpy().astype(np.int64)


def evaluate_sharded(gt_lists, pred_lists, thresholds=batch.COCO_THRESHOLDS, group=None, gather=True,
                     compute_fn=None):
    """Dataset-level detection counts over all ranks.

    gt_lists / pred_lists: per image, lists of COCO RLE dicts (every rank passes the full
    dataset description; only its shard is touched).  Returns a dict with
      'totals'    int64 [T, 3]  TP/FP/FN summed over the dataset (identical on every rank)
      'per_image' int64 [n_images, T, 3] in image order (if gather) else the local shard
      'index'     the image indices this rank evaluated
    compute_fn(gt_shard, pred_shard, thresholds) -> [n_local, T, 3] overrides the GPU pipeline
    (used by the CPU tests)."""
    rank, world = world_info(group)
    n = len(gt_lists)
    idx = shard_indices(n, rank, world)
    fn = compute_fn or _gpu_match_counts
    local = np.asarray(fn([gt_lists[i] for i in idx], [pred_lists[i] for i in idx], thresholds), np.int64)
    T = len(thresholds)
    local = local.reshape(len(idx), T, 3)
    totals = torch.from_numpy(local.sum(axis=0).reshape(-1).copy())
    use_cuda = world > 1 and dist.get_backend(group) == 'nccl'
    if use_cuda:
        totals = totals.cuda()
    all_reduce_sum_(totals, group)
    out = {'totals': totals.cpu().numpy().reshape(T, 3), 'index': idx}
    if gather:
        out['per_image'] = gather_rows(local.reshape(len(idx), -1), idx, n, group).reshape(n, T, 3)
    else:
        out['per_image'] = local
    return out


def _gpu_satellite_counts(part_lists, sat_lists, thresh, n_bins):
    n = len(part_lists)
    if n == 0:
        return np.zeros((0, 4), np.int64), np.zeros(n_bins, np.int64)
    masks = []
    for p, s in zip(part_lists, sat_lists):
        masks += list(s) + list(p)
    table = engine.table_from_rle(masks, layout=engine.LAYOUT_CROP)
    groups = engine.Groups.interleaved(table.device, [len(s) for s in sat_lists], [len(p) for p in part_lists])
    rows = engine.intersect_rows(table, groups, engine.MODE_SAT)
    counts, hist = engine.satellite_counts(table, rows, groups, thresh, n_bins)
    return counts.cpu().numpy().astype(np.int64), hist.cpu().numpy()


def satellites_sharded(particle_lists, satellite_lists, thresh=0.5, n_bins=64, group=None, compute_fn=None):
    """Dataset-level satellite statistics over all ranks (the sums of powder.py:525-547).

    Returns totals dict {n_images, n_particles, n_satellites, n_satellites_unmatched,
    n_satellited_particles, sat_frac} and the satellites-per-particle histogram (bin b counts
    particles owning b satellites, last bin clamps) -- both identical on every rank."""
    rank, world = world_info(group)
    n = len(particle_lists)
    idx = shard_indices(n, rank, world)
    fn = compute_fn or _gpu_satellite_counts
    counts, hist = fn([particle_lists[i] for i in idx], [satellite_lists[i] for i in idx], thresh, n_bins)
    counts = np.asarray(counts, np.int64).reshape(len(idx), 4)
    payload = torch.from_numpy(np.concatenate([[len(idx)], counts.sum(axis=0), np.asarray(hist, np.int64)]))
    if world > 1 and dist.get_backend(group) == 'nccl':
        payload = payload.cuda()
    all_reduce_sum_(payload, group)
    p = payload.cpu().numpy()
    n_img, matched, unmatched, sat_particles, particles = p[:5]
    return {'n_images': int(n_img), 'n_particles': int(particles), 'n_satellites': int(matched),
            'n_satellites_unmatched': int(unmatched), 'n_satellited_particles': int(sat_particles),
            'sat_frac': sat_particles / particles if particles else float('nan'),
            'spp_hist': p[5:].copy(), 'index': idx}


def area_histogram_sharded(mask_lists, lo, bin_width, n_bins, group=None, compute_fn=None):
    """Binned mask-area histogram of the dataset (size-distribution payload), summed over ranks."""
    rank, world = world_info(group)
    idx = shard_indices(len(mask_lists), rank, world)

    def gpu(shard):
        masks = [m for ml in shard for m in ml]
        if not masks:
            return np.zeros(n_bins, np.int64)
        t = engine.table_from_rle(masks, paint=False)
        return engine.hist_u32(t.area[:t.n], lo, bin_width, n_bins).cpu().numpy()

    hist = torch.from_numpy(np.asarray((compute_fn or gpu)([mask_lists[i] for i in idx]), np.int64).copy())
    if world > 1 and dist.get_backend(group) == 'nccl':
        hist = hist.cuda()
    all_reduce_sum_(hist, group)
    return hist.cpu().numpy()
