#!/usr/bin/env python
"""Host <-> device copy rate of this box from pinned memory, with and without pinning the process to the CPUs
NVML reports as local to the GPU (NUMA placement of the pinned pages): python profiles/h2d_probe.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def rate(nbytes, h2d=True, reps=8):
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)                                   # first touch on the current CPU set
    dev = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
    for _ in range(2):
        (dev.copy_(host, non_blocking=True) if h2d else host.copy_(dev, non_blocking=True))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        (dev.copy_(host, non_blocking=True) if h2d else host.copy_(dev, non_blocking=True))
    e1.record()
    torch.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) / 1e3) / 1e9


def main():
    from bench import gpu_local_cpus
    torch.cuda.set_device(0)
    out = {'cpus_allowed_before': len(os.sched_getaffinity(0))}
    for n in (8 << 20, 96 << 20, 512 << 20):
        out['h2d_GBs_%dMB' % (n >> 20)] = rate(n, True)
        out['d2h_GBs_%dMB' % (n >> 20)] = rate(n, False)
    cpus = gpu_local_cpus(0)
    out['gpu_local_cpus'] = len(cpus) if cpus else None
    if cpus:
        os.sched_setaffinity(0, cpus)
        for n in (96 << 20, 512 << 20):
            out['h2d_GBs_%dMB_local' % (n >> 20)] = rate(n, True)
            out['d2h_GBs_%dMB_local' % (n >> 20)] = rate(n, False)
    try:
        import pynvml
        pynvml.nvmlInit()
        hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
        out['pcie_gen_cur_max'] = [pynvml.nvmlDeviceGetCurrPcieLinkGeneration(hnd), pynvml.nvmlDeviceGetMaxPcieLinkGeneration(hnd)]
        out['pcie_width_cur_max'] = [pynvml.nvmlDeviceGetCurrPcieLinkWidth(hnd), pynvml.nvmlDeviceGetMaxPcieLinkWidth(hnd)]
    except Exception as e:
        out['pcie'] = 'nvml: %s' % e
    print(json.dumps(out))


if __name__ == '__main__':
    main()
