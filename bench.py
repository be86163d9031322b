#!/usr/bin/env python
"""bench.py -- throughput of the mask-evaluation hot path on B200 (see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--layout full|span] [--config NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU algorithm on the host cores

A step is one pass of the hot path over one batch of synthetic images (BASELINE.json
configs[1]: 1,000 images of 1024x1024 with 500 GT x 500 predicted masks each, per GPU):
RLE run counts resident in HBM -> per-mask measurements -> bit-packed masks -> bbox-pruned
intersections (dense int32 G x P matrix out) + per-GT arg-max IoU -> TP/FP/FN at IoU
0.50:0.05:0.95 per image and in total (+ one NCCL all-reduce of the totals when N > 1).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'mask-pair IoUs/sec (GT x pred mask pairs matched and scored, IoU 0.50:0.95)'
UNIT = 'pairs/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c2_powder_batch')
    ap.add_argument('--images', type=int, default=1000, help='images per GPU per step')
    ap.add_argument('--layout', default='full', choices=['full', 'span', 'crop'],
                    help='full = canonical full-frame packed masks (the roofline accounting of SURVEY 8d); '
                         'span = culled storage (only first..last 1-pixel of each mask); '
                         'crop = bounding-box windows (the cropped accounting of SURVEY 8d)')
    ap.add_argument('--sub', type=int, default=0, help='images per launch group (0 = auto)')
    ap.add_argument('--kernel', default='rows', choices=['rows', 'mma', 'mma2', 'grid', 'scan'],
                    help='intersection kernel: rows = bbox-culled AND+popc (default; crop layout with >= 1024 '
                         'columns per image prunes through a uniform grid); grid / scan = crop layout with the grid '
                         'forced / forbidden; mma = dense int8 tcgen05 contraction (for crowded images, e.g. '
                         '--config dense_overlap)')
    ap.add_argument('--sparse', action='store_true',
                    help='crop layout: no dense G x P matrix, the non-zero intersections come out as triplets '
                         '(bbox-pruned sparse IoU, the C4 form of SURVEY 8d)')
    ap.add_argument('--mma-sort', action='store_true',
                    help='with --kernel mma: cut the tiles from spatially sorted masks (fewer slabs contracted; the '
                         'default contracts the full pixel range so that the tensor roofline counts executed work)')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--graph', action='store_true', help='replay each step from a CUDA graph')
    ap.add_argument('--span-sub', type=int, default=0, help='images per launch group of the span run')
    ap.add_argument('--unfused', action='store_true', help='separate measure / scan / paint launches')
    ap.add_argument('--no-span', action='store_true', help='skip the secondary span-layout measurement')
    ap.add_argument('--cpu-images', type=int, default=0)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


def tensor_peak_tops():
    """int8 tensor peak in TOP/s: twice the measured dense bf16 rate (the i8 pipe is 2x bf16 on sm_100;
    nominal 4.5 POP/s).  Burst figure: the contraction draws ~230 W and holds 1965 MHz, it never
    reaches the power cap that defines the sustained bf16 number."""
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return 2.0 * float(d['bf16_tflops']), '2 x measured bf16 burst (MEASURED_PEAKS.json bf16_tflops)'
    return 2.0 * 1590.0, '2 x fallback bf16 burst (B200_PROFILING.md)'


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw'

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(',')]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'note': 'nvidia-smi unavailable'}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 and len(r) >= 6] or [r for _, r in self.rows if len(r) >= 6]
        if not rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'note': 'no samples'}
        sm = sorted(float(r[0]) for r in rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith('active') for r in rows)]
        out = {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(rows[0][1]), 'reasons': reasons, 'samples': len(rows)}
        try:
            out['power_w_max'] = max(float(r[6]) for r in rows)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's algorithm (oracle port) on the host cores
# ------------------------------------------------------------------------------------------
def _cpu_task(args):
    gt, pr, th, mode = args
    from oracle import ampis_ref as R
    if mode == 0:
        r = R.det_seg_scores(gt, pr, th)
        return len(r['det_tp']), len(r['det_fp']), len(r['det_fn'])
    r = R.rle_satellite_match(pr, gt, th)
    return len(r['satellite_matches']), 0, len(r['satellites_unmatched'])


def cpu_images(cfg_name, n_img, seed):
    """n_img synthetic images as lists of compressed-RLE dicts (what the reference consumes)."""
    from ampis_b200 import batch
    from oracle import cocomask as rle
    host = batch.synth(cfg_name, n_img, seed)
    out = []
    for g in range(n_img):
        rows, cols = host.image_masks(g)
        size = [host.h, host.w]
        out.append(([{'size': size, 'counts': rle.string_from_counts(c)} for c in rows],
                    [{'size': size, 'counts': rle.string_from_counts(c)} for c in cols]))
    return host, out


def cpu_run(pool, images, thresholds, mode):
    tasks = [(gt, pr, float(t), mode) for gt, pr in images for t in thresholds]
    t0 = time.perf_counter()
    res = pool.map(_cpu_task, tasks, chunksize=1)
    return time.perf_counter() - t0, res


def reference_arm(args):
    """--impl reference: per step, `n_img` images of the same synthetic workload through the
    reference's loops (analyze.py:149-172 + 315-327 restated in oracle/ampis_ref.py over the C
    restatement of pycocotools), one det_seg_scores call per IoU threshold as a user of the
    reference would do, on all host cores.  kind = "port": pycocotools itself is not installable
    here."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    from ampis_b200 import batch
    from oracle import cocomask
    cocomask.build()
    cores = os.cpu_count() or 1
    cfg = batch.CONFIGS[args.config]
    thresholds = batch.COCO_THRESHOLDS if cfg['mode'] == 0 else [0.5]
    per_task = 0.75 if cfg['mode'] == 0 else 2.5        # seconds per (image, threshold) on one core, measured
    budget = min(8.0, 200.0 / max(args.steps + args.warmup, 1))
    n_img = args.cpu_images or max(1, int(cores * budget / (per_task * len(thresholds))))
    host, images = cpu_images(args.config, n_img, 777)
    pairs = n_img * host.n_rows * host.n_cols
    with mp.get_context('fork').Pool(cores) as pool:
        for _ in range(args.warmup):
            cpu_run(pool, images[:max(1, min(n_img, cores // len(thresholds)))], thresholds, cfg['mode'])
        t = 0.0
        for _ in range(args.steps):
            dt, _ = cpu_run(pool, images, thresholds, cfg['mode'])
            t += dt
    value = pairs * args.steps / t
    sample = '%d synthetic %s images per step, %d IoU thresholds each, oracle port of the reference loops, ' \
             '%d processes' % (n_img, args.config, len(thresholds), cores)
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * t / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u32', 'data': 'synthetic',
        'images_per_s': n_img * args.steps / t,
        'config': {'workload': workload_name(args, host), 'images_per_step': n_img},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def workload_name(args, host):
    return '%s: %dx%d px, %d x %d masks per image' % (args.config, host.w, host.h, host.n_rows, host.n_cols)


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
KERNELS = ['measure+scan', 'paint', 'rows', 'counts']


class LayoutRun(object):
    """Inputs of one rank resident in HBM + the workspace for one storage layout."""

    def __init__(self, args, dev, rank, layout, sub):
        import torch
        from ampis_b200 import batch, engine
        self.layout, self.sub, self.dev = layout, sub, dev
        self.fused = not args.unfused
        self.last_table = None
        t0 = time.time()
        self.subs = []
        for s0 in range(0, args.images, sub):
            k = min(sub, args.images - s0)
            host = batch.synth(args.config, k, 1_000_003 * (rank + 1) + s0)
            self.subs.append(batch.DeviceBatch(host, dev, dense=not (args.sparse and layout == engine.LAYOUT_CROP)))
        self.t_gen = time.time() - t0
        self.total_runs = sum(b.host.total_runs() for b in self.subs)
        need = [batch.arena_chunks_needed(b, layout) for b in self.subs]
        self.stored_chunks = sum(need)
        self.arena = torch.empty(4 * max(need), dtype=torch.int32, device=dev)
        n_rows = max(b.groups.n_rows for b in self.subs)
        self.rows_out = engine.RowResult(torch.empty(n_rows, dtype=torch.int32, device=dev),
                                         torch.empty(n_rows, dtype=torch.int32, device=dev),
                                         torch.empty(n_rows, dtype=torch.float64, device=dev),
                                         torch.empty(max(max(b.groups.imat_size for b in self.subs), 1),
                                                     dtype=torch.int32, device=dev))
        self.thresholds = batch.COCO_THRESHOLDS
        cfg = batch.CONFIGS[args.config]
        n_tot = len(self.thresholds) * 3
        # one int64 payload for the all-reduce: TP/FP/FN x thresholds [+ binned area histogram]
        self.payload = torch.zeros(n_tot + cfg.get('area_bins', 0), dtype=torch.int64, device=dev)
        self.totals = self.payload[:n_tot]
        self.area_hist = self.payload[n_tot:] if cfg.get('area_bins') else None
        self.kernel = args.kernel
        self.mode = cfg['mode']
        self.pipes = [batch.Pipeline(b, layout, self.arena, self.rows_out, self.thresholds, self.totals,
                                     fused=self.fused, kernel=args.kernel, mma_sort=args.mma_sort,
                                     area_hist=self.area_hist, area_bin_width=cfg.get('area_bin_width', 64),
                                     sparse_capacity=(32 * b.groups.n_rows if args.sparse and
                                                      layout == engine.LAYOUT_CROP else None)) for b in self.subs]
        if self.mode != 0:
            for p in self.pipes[1:]:
                p.spp_hist = self.pipes[0].spp_hist
        self.graph = None

    def launch_all(self, record=None):
        import torch
        self.payload.zero_()
        if self.mode != 0:
            self.pipes[0].spp_hist.zero_()
        for p in self.pipes:
            if record is None:
                p.launch()
            else:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
                p.launch(lambda i, ev=ev: ev[i].record())
                record.append(ev)

    def capture(self):
        """Capture one step's launch sequence in a CUDA graph (launch-bound for small launch groups)."""
        import torch
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.launch_all()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.launch_all()

    def step(self, world, dist, record=None):
        """one pass over the rank's batch; `record` collects the CUDA events around each kernel group"""
        if self.graph is not None and record is None:
            self.graph.replay()
        else:
            self.launch_all(record)
        if world > 1:       # TP/FP/FN x thresholds (or the satellites-per-particle histogram): the only exchange
            dist.all_reduce(self.payload if self.mode == 0 else self.pipes[0].spp_hist)
        return self.totals

    def timed(self, args, world, dist, sync):
        import torch
        for _ in range(max(args.warmup, 3)):
            self.step(world, dist)
        sync()
        record = []
        use_graph = self.graph is not None
        if use_graph:     # per-kernel shares come from an instrumented pass outside the timed region
            for _ in range(2):
                self.launch_all(record)
            sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            self.step(world, dist, None if use_graph else record)
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=self.dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        kt = np.zeros(4)
        for ev in record:
            for i in range(4):
                kt[i] += ev[i].elapsed_time(ev[i + 1])
        self.pipes[-1].table.check()      # arena large enough, RLE well-formed (after the timed region)
        self.sparse_pairs = None
        if self.pipes[0].sparse is not None:
            cnt = [int(p.sparse.count.item()) for p in self.pipes]
            assert all(c <= p.sparse.capacity for c, p in zip(cnt, self.pipes)), 'sparse triplet list overflowed'
            self.sparse_pairs = sum(cnt)
        for p in self.pipes:
            if getattr(p.grid, 'capacity', None):
                assert p.grid.needed() <= p.grid.capacity, 'grid entry list overflowed'
        if self.mode != 0:      # satellites: per-image counts summed over the batch + the global histogram
            c = sum(p.counts.cpu().numpy().sum(axis=0) for p in self.pipes)
            return float(ms.item()), kt, np.concatenate([c, self.pipes[0].spp_hist.cpu().numpy()[:8]]).reshape(1, -1)
        return float(ms.item()), kt, self.totals.cpu().numpy().reshape(-1, 3)


def roofline_of(args, cfg, run, ms, kt, world):
    """Roofline of the dominant kernel + the whole step against the canonical accounting."""
    from ampis_b200 import engine
    peak, peak_src = peaks()
    n_img = args.images
    per_image = cfg['n_rows'] + cfg['n_cols']
    B_m = ((cfg['h'] * cfg['w'] + 127) // 128) * 16
    n_masks = n_img * per_image
    pairs_img = cfg['n_rows'] * cfg['n_cols']
    # SURVEY 8d per-kernel figures: decode = 4R + bytes stored; intersection = (G+P)*B_m + 4*G*P
    alg = {'paint': 4 * run.total_runs + run.stored_chunks * 16, 'rows': n_masks * B_m + 4 * n_img * pairs_img}
    canonical_img = 4 * run.total_runs / n_img + 2 * per_image * B_m + 4 * pairs_img + 8 * per_image + 32 * cfg['n_rows']
    dom = 'paint' if kt[1] >= kt[2] else 'rows'
    launches = len(run.subs)
    dur_ms = kt[KERNELS.index(dom)] / ((2 if run.graph is not None else args.steps) * launches)
    achieved = alg[dom] / launches / (dur_ms / 1e3) / 1e9
    lay = {engine.LAYOUT_FULL: 'full', engine.LAYOUT_SPAN: 'span', engine.LAYOUT_CROP: 'crop'}[run.layout]
    # DRAM bytes per launch of the dominant kernel (ncu dram__bytes_read.sum + dram__bytes_write.sum of the
    # capture summarised in profiles/), scaled from the captured images per launch to this run's
    traffic, traffic_detail = None, None
    tp = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tp):
        traffic_detail = json.load(open(tp)).get('%s/%s/%s' % (args.config, lay, dom))
        if traffic_detail:
            traffic = traffic_detail['bytes_per_image'] * n_img / launches
    step_gbs = canonical_img * n_img * args.steps / (ms / 1e3) / 1e9
    if run.kernel in ('mma', 'mma2') and dom == 'rows':
        # dense contraction: 2*G*P*H*W integer ops per image, tensor-pipe bound
        tpeak, tsrc = tensor_peak_tops()
        ops = 2.0 * pairs_img * cfg['h'] * cfg['w'] * n_img / launches
        ach = ops / (dur_ms / 1e3) / 1e12
        return {'bound': 'tensor', 'kernel': 'intersect_mma_kernel (+ rows_from_imat_kernel)', 'achieved': ach,
                'peak': tpeak, 'unit': 'TOP/s', 'frac': ach / tpeak, 'traffic': None, 'peak_source': tsrc,
                'algorithmic_ops_per_launch': ops, 'launch_ms': dur_ms,
                'step_canonical': {'bytes_per_image': canonical_img, 'achieved': step_gbs, 'frac': step_gbs / peak},
                'kernel_share': {k: float(v) for k, v in zip(KERNELS, kt / kt.sum())}}
    pk = 'rle_paint_kernel' if args.unfused else 'rle_measure_paint_kernel'
    rk = 'intersect_rows_kernel'
    if run.layout == engine.LAYOUT_CROP:
        rk = 'intersect_rows_grid_kernel' if getattr(run.pipes[0].grid, 'capacity', None) else 'intersect_rows_crop_kernel'
        if not args.unfused and 0 < run.total_runs / max(n_masks, 1) <= 112:
            pk = 'rle_measure_paint_crop_kernel'       # 8 or 16 lanes per mask
    return {'bound': 'hbm', 'kernel': {'paint': pk, 'rows': rk}[dom],
            'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
            'traffic_detail': traffic_detail,
            'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg[dom] / launches, 'launch_ms': dur_ms,
            'peak_note': 'the measured peak is a device-to-device COPY (half reads, half writes); the full-frame decode '
                         'kernel is a pure write stream (4R + stored bytes, 0.3 % reads) and sustains slightly more '
                         'than the copy figure, hence frac a little above 1 (ncu: dram write 6.9 TB/s)' if dom == 'paint'
                         and run.layout == engine.LAYOUT_FULL else None,
            'step_canonical': {'bytes_per_image': canonical_img, 'achieved': step_gbs, 'frac': step_gbs / peak,
                               'note': 'whole step per GPU vs the two-pass full-frame accounting of SURVEY 8d; '
                                       'bbox/span culling lets the step move fewer bytes than that'},
            'kernel_share': {k: float(v) for k, v in zip(KERNELS, kt / kt.sum())}}


def cropped_accounting(cfg, run, ms, args):
    """SURVEY 8d's accounting for culled storage (prescribed for C4): bytes_img = 4 R_img + 2 sum(window bytes)
    + 24 N_pairs + 8 (G + P) + 32 G, window bytes as this layout stores them (32-row bands of the box columns),
    N_pairs = the non-zero intersections when the run produced the sparse list (a lower bound of the
    bbox-overlapping pairs), else the dense 4 G P matrix."""
    peak, peak_src = peaks()
    n_img = args.images
    per_image = cfg['n_rows'] + cfg['n_cols']
    pairs_term = 24.0 * run.sparse_pairs / n_img if run.sparse_pairs is not None else 4.0 * cfg['n_rows'] * cfg['n_cols']
    bytes_img = 4.0 * run.total_runs / n_img + 2.0 * run.stored_chunks * 16 / n_img + pairs_term + 8 * per_image + \
        32 * cfg['n_rows']
    gbs = bytes_img * n_img * args.steps / (ms / 1e3) / 1e9
    return {'bytes_per_image': bytes_img, 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak,
            'output': 'sparse triplets' if run.sparse_pairs is not None else 'dense G x P matrix',
            'note': 'the culled step is bound by instruction issue / load latency, not by these bytes (DESIGN 8)'}


def main():
    args = parse()
    if args.impl == 'reference':
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from ampis_b200 import batch, engine
    from ampis_b200 import _native as N

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    N.lib()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = batch.CONFIGS[args.config]
    strong = 'dataset_images' in cfg
    if strong:          # a fixed dataset split over the ranks (strong scaling); every rank gets its share
        total_images = cfg['dataset_images']
        args.images = total_images // world + (1 if rank < total_images % world else 0)
    job_images = cfg['dataset_images'] if strong else world * args.images        # images all ranks evaluate per step
    layout = {'full': engine.LAYOUT_FULL, 'span': engine.LAYOUT_SPAN, 'crop': engine.LAYOUT_CROP}[args.layout]
    per_image = cfg['n_rows'] + cfg['n_cols']
    B_m = ((cfg['h'] * cfg['w'] + 127) // 128) * 16

    def auto_sub(lay):
        if args.sub:
            return args.sub
        if lay == engine.LAYOUT_FULL:
            by_imat = max(1, int(4e9 // (4 * cfg['n_rows'] * cfg['n_cols'])))
            return max(1, min(args.images, int(12e9 // (per_image * B_m)), by_imat))       # ~12 GB arena
        by_imat = max(1, int(4e9 // (4 * cfg['n_rows'] * cfg['n_cols'])))          # dense matrices <= 4 GB per launch
        if args.sparse and lay == engine.LAYOUT_CROP:
            by_imat = args.images
        # culled layouts: the arena is small, so all images of the step go out in ONE launch of each kernel (bounded by
        # 4 GB of dense matrices) -- at 91 or 250 images per launch the short kernels of this step were mostly ramp and
        # tail (profiles/kernels_r02.md: 2.69 / 1.83 / 1.45 ms per 1,000 C2 images at 91 / 250 / 1,000 per launch)
        return min(args.images, by_imat)

    sampler = ClockSampler(local) if rank == 0 else None       # samples cover warm-up + timed steps
    wall0 = time.time()
    run = LayoutRun(args, dev, rank, layout, auto_sub(layout))
    if args.graph:
        run.capture()
    ms, kt, final_totals = run.timed(args, world, dist, sync)
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if sampler else None

    # ---- end to end through host buffers: compressed RLE strings in pinned memory -> H2D -> GPU string
    # decode -> same pipeline -> D2H of per-image counts and per-GT matches
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, run.subs, dev, layout, run.arena, run.rows_out, run.thresholds, world, dist, sync,
                      pipes=run.pipes)

    # ---- the culled storage layout (product default), measured beside the canonical one
    span = None
    if layout == engine.LAYOUT_FULL and not args.no_span:
        del run.arena
        srun = LayoutRun(args, dev, rank, engine.LAYOUT_SPAN, args.span_sub or auto_sub(engine.LAYOUT_SPAN))
        if args.graph:
            srun.capture()
        sms, skt, stot = srun.timed(args, world, dist, sync)
        assert np.array_equal(stot, final_totals), 'span and full layouts disagree'
        span = {'value': job_images * cfg['n_rows'] * cfg['n_cols'] * args.steps / (sms / 1e3), 'unit': UNIT,
                'images_per_s': job_images * args.steps / (sms / 1e3), 'ms_per_step': sms / args.steps,
                'images_per_launch': srun.sub, 'roofline': roofline_of(args, cfg, srun, sms, skt, world),
                'note': 'same inputs and bit-identical results; only first..last 1-pixel of each mask is stored'}
        if not args.no_e2e:
            span['e2e'] = run_e2e(args, srun.subs, dev, engine.LAYOUT_SPAN, srun.arena, srun.rows_out,
                                  srun.thresholds, world, dist, sync)
    # ---- bounding-box windows: the smallest storage, same results
    crop = None
    if layout == engine.LAYOUT_FULL and not args.no_span and args.kernel == 'rows':
        del srun.arena
        crun = LayoutRun(args, dev, rank, engine.LAYOUT_CROP, args.span_sub or auto_sub(engine.LAYOUT_CROP))
        if args.graph:
            crun.capture()
        cms, ckt, ctot = crun.timed(args, world, dist, sync)
        assert np.array_equal(ctot, final_totals), 'crop and full layouts disagree'
        crop = {'value': job_images * cfg['n_rows'] * cfg['n_cols'] * args.steps / (cms / 1e3), 'unit': UNIT,
                'images_per_s': job_images * args.steps / (cms / 1e3), 'ms_per_step': cms / args.steps,
                'images_per_launch': crun.sub, 'roofline': roofline_of(args, cfg, crun, cms, ckt, world),
                'stored_bytes_per_image': crun.stored_chunks * 16 / args.images,
                'intersection_kernel': 'grid' if getattr(crun.pipes[0].grid, 'capacity', None) else 'scan',
                'note': 'same inputs and bit-identical results; only the bounding-box window of each mask is '
                        'stored (32-row bands of the box columns)'}
        crop['cropped_accounting'] = cropped_accounting(cfg, crun, cms, args)
        if crun.sparse_pairs is not None:
            crop['nonzero_pairs_per_image'] = crun.sparse_pairs / args.images
        if not args.no_e2e:
            crop['e2e'] = run_e2e(args, crun.subs, dev, engine.LAYOUT_CROP, crun.arena, crun.rows_out,
                                  crun.thresholds, world, dist, sync, pipes=crun.pipes)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    host0 = run.subs[0].host
    n_masks = args.images * per_image
    value = job_images * cfg['n_rows'] * cfg['n_cols'] * args.steps / (ms / 1e3)
    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong' if strong else 'weak',
        'vs_baseline': None, 'dtype': 'u32', 'data': 'synthetic',
        'images_per_s': job_images * args.steps / (ms / 1e3),
        'config': {'workload': workload_name(args, host0), 'images_per_gpu_per_step': args.images,
                   'layout': args.layout, 'sparse_output': bool(run.pipes[0].sparse), 'intersection_kernel': ('grid' if getattr(run.pipes[0].grid, 'capacity', None) else args.kernel) + ('+sorted tiles' if args.kernel in ('mma', 'mma2') and args.mma_sort else ''), 'images_per_launch': run.sub, 'thresholds': 'IoU 0.50:0.05:0.95' if cfg['mode'] == 0 else 'satellite overlap > 0.5',
                   'runs_per_mask': run.total_runs / n_masks,
                   'l2': 'per step the kernels stream %.0f MB of run counts and a %.1f GB packed-mask arena, both '
                         'larger than the 126 MB L2; no explicit flush' % (4 * run.total_runs / 1e6,
                                                                          run.stored_chunks * 16 / len(run.subs) / 1e9),
                   'parallelism': 'images sharded over %d GPU(s); one int64 all-reduce of TP/FP/FN%s per step' % (
                       world, ' + %d-bin area histogram' % cfg['area_bins'] if cfg.get('area_bins') else '')},
        'roofline': roofline_of(args, cfg, run, ms, kt, world),
        # grid-pruned crop rows: + the grid build kernel
        'gpu_launches': int(args.steps * len(run.subs) * ((7 if args.unfused else 3) + (1 if args.kernel in ('mma', 'mma2') else 0) +
                                                            (1 if getattr(run.pipes[0].grid, 'capacity', None) else 0) +
                                                            (1 if cfg.get('area_bins') else 0))),
        ('totals_tp_fp_fn_at_0.50' if cfg['mode'] == 0 else 'sat_matched_unmatched_satellited_particles+spp_hist'): final_totals[0].tolist(),
        'clocks': clocks, 'setup_s': {'synthesize+upload': run.t_gen},
    }
    if run.sparse_pairs is not None:
        out['config']['nonzero_pairs_per_image'] = run.sparse_pairs / args.images
    if layout == engine.LAYOUT_CROP:
        out['cropped_accounting'] = cropped_accounting(cfg, run, ms, args)
    if e2e:
        out['e2e'] = e2e
    if span:
        out['span_layout'] = span
    if crop:
        out['crop_layout'] = crop
    if not args.no_cpu:
        out['cpu_baseline'] = cpu_baseline(args)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, subs, dev, layout, arena, rows_out, thresholds, world, dist, sync, pipes=None):
    """Same metric through host buffers: every step copies the compressed RLE strings (what the
    reference API receives) from pinned host memory, decodes them on the GPU, runs the pipeline
    and reads the per-image counts and per-GT matches back."""
    import torch
    from ampis_b200 import engine
    from ampis_b200 import _native as N
    _p, _s = engine._p, engine._stream
    pinned = []
    for b in subs:       # setup (untimed): strings produced by the GPU encoder, parked in pinned memory
        n = b.host.n_masks
        lens = b.cnt_len.cpu().numpy().astype(np.int64)
        choff = np.zeros(n + 1, np.int64)
        np.cumsum(7 * lens, out=choff[1:])
        d_choff = torch.from_numpy(choff).to(dev)
        chars = torch.empty(int(choff[-1]), dtype=torch.uint8, device=dev)
        chlen = torch.empty(n, dtype=torch.int32, device=dev)
        N.call('ampis_rle_string_encode', _p(b.cnt), _p(b.cnt_off), _p(b.cnt_len), n, _p(chars), _p(d_choff),
               _p(chlen), _s())
        ln = chlen.cpu().numpy().astype(np.int64)
        buf = chars.cpu().numpy()
        off = np.zeros(n + 1, np.int64)
        np.cumsum(ln, out=off[1:])
        blob = np.concatenate([buf[choff[i]:choff[i] + ln[i]] for i in range(n)]) if n else np.zeros(0, np.uint8)
        pinned.append((torch.from_numpy(blob).pin_memory(), torch.from_numpy(off).pin_memory(), b))
    max_chars = max(p[0].numel() for p in pinned)
    d_chars = torch.empty(max_chars, dtype=torch.uint8, device=dev)
    d_cnt = torch.empty(max_chars, dtype=torch.int32, device=dev)
    n_thr = len(thresholds)
    h_out = []
    for blob, off, b in pinned:
        shape = (b.groups.n_groups, n_thr, 3) if b.mode == engine.MODE_IOU else (b.groups.n_groups, 4)
        h_out.append((torch.empty(shape, dtype=torch.int32).pin_memory(),
                      torch.empty(b.groups.n_rows, dtype=torch.int32).pin_memory(),
                      torch.empty(b.groups.n_rows, dtype=torch.float64).pin_memory()))
    h2d = sum(p[0].numel() + p[1].numel() * 8 for p in pinned)
    d2h = sum(o[0].numel() * 4 + o[1].numel() * 4 + o[2].numel() * 8 for o in h_out)
    totals = torch.zeros(n_thr * 3, dtype=torch.int64, device=dev)
    spp_hist = torch.zeros(64, dtype=torch.int64, device=dev)

    # the strings of launch group i+1 are uploaded on a copy stream while group i is evaluated:
    # two device buffers, handed back and forth with events
    copy_stream = torch.cuda.Stream(device=dev)
    max_off = max(p[1].numel() for p in pinned)
    bufs = [(d_chars, torch.empty(max_off, dtype=torch.int64, device=dev)),
            (torch.empty_like(d_chars), torch.empty(max_off, dtype=torch.int64, device=dev))]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    max_masks = max(b.host.n_masks for b in subs)
    d_cnt_len = torch.empty(max_masks, dtype=torch.int32, device=dev)

    def step():
        totals.zero_()
        spp_hist.zero_()
        comp = torch.cuda.current_stream()
        copy_stream.wait_stream(comp)                 # uploads of this step start inside the timed region
        for i, ((blob, off, b), (hc, hb, hs)) in enumerate(zip(pinned, h_out)):
            n = b.host.n_masks
            k = i & 1
            dc, d_off = bufs[k][0][:blob.numel()], bufs[k][1][:off.numel()]
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_free[k])    # the evaluation that last read this buffer has finished
                dc.copy_(blob, non_blocking=True)
                d_off.copy_(off, non_blocking=True)
                ev_ready[k].record(copy_stream)
            comp.wait_event(ev_ready[k])
            cnt_len = d_cnt_len[:n]
            N.call('ampis_rle_string_decode', _p(dc), _p(d_off), n, _p(d_cnt), _p(d_off), _p(cnt_len), _s())
            t = engine.MaskTable(dev, n, d_cnt, d_off, cnt_len, b.h, b.w, layout)
            if args.unfused:
                t.measure().paint(arena)
            else:
                t.measure_paint(arena)
            if args.kernel in ('mma', 'mma2'):
                rows = engine.intersect_mma(t, b.groups, b.mode, out=rows_out, sort=args.mma_sort, pair=args.kernel == 'mma2')
            else:
                rows = engine.intersect_rows(t, b.groups, b.mode, out=rows_out, grid=pipes[i].grid if pipes else None,
                                             sparse=pipes[i].sparse if pipes else None,
                                             pairs=pipes[i].pairs if pipes else None)
            if b.mode == engine.MODE_IOU:
                counts, _ = engine.match_counts(rows, b.groups, thresholds, totals=totals)
            else:       # satellites: per-image (matched, unmatched, satellited particles, particles) + global histogram
                counts, _ = engine.satellite_counts(t, rows, b.groups, 0.5, hist=spp_hist)
            hc.copy_(counts, non_blocking=True)
            hb.copy_(rows.best_col[:b.groups.n_rows], non_blocking=True)
            hs.copy_(rows.best_score[:b.groups.n_rows], non_blocking=True)
            ev_free[k].record(comp)
        red = totals if subs[0].mode == engine.MODE_IOU else spp_hist
        if world > 1:
            dist.all_reduce(red)
        return red.cpu()

    for _ in range(2):
        step()
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    sync()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    n_img = sum(b.host.n_images for b in subs)
    pairs = world * n_img * subs[0].host.n_rows * subs[0].host.n_cols
    return {'value': pairs * args.steps / (ms / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d),
            'd2h_bytes_per_step': int(d2h), 'ms_per_step': ms / args.steps,
            'input': 'COCO-compressed RLE strings in pinned host memory, uploaded on a copy stream (double-buffered) '
                     'and decoded on the GPU'}


def cpu_baseline(args):
    """The reference's CPU algorithm (oracle port) on a bounded sample of the same workload, run in a
    fresh process (no CUDA context is forked): one step of the --impl reference arm sized to ~15 s."""
    cores = os.cpu_count() or 1
    from ampis_b200 import batch
    cfg = batch.CONFIGS[args.config]
    n_thr = len(batch.COCO_THRESHOLDS) if cfg['mode'] == 0 else 1
    per_task = 0.75 if cfg['mode'] == 0 else 2.5
    n_img = args.cpu_images or max(1, int(cores * 15.0 / (per_task * n_thr)))
    env = dict(os.environ, RANK='0', WORLD_SIZE='1', CUDA_VISIBLE_DEVICES='')
    r = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '1',
                        '--warmup', '0', '--config', args.config, '--cpu-images', str(n_img)],
                       env=env, capture_output=True, text=True, timeout=900)
    line = [l for l in r.stdout.splitlines() if l.startswith('{')]
    if r.returncode != 0 or not line:
        return {'value': None, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': 'failed: ' + r.stderr[-300:]}
    d = json.loads(line[-1])
    cb = d['cpu_baseline']
    cb['images_per_s'] = d['images_per_s']
    return cb


if __name__ == '__main__':
    main()
