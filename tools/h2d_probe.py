#!/usr/bin/env python
"""Concurrent pinned-host -> device copy bandwidth on every GPU of the box, with and without binding each
process to the CPUs NVML reports as local to its GPU.  Launch under torchrun (one rank per GPU):
   python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_probe.py [--bind]"""
import argparse
import os
import time


def gpu_cpus(index):
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(index)
    ncpu = os.cpu_count()
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
    cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1]
    return [c for c in cpus if c < ncpu]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bind", action="store_true")
    ap.add_argument("--gb", type=float, default=2.0)
    a = ap.parse_args()
    rank = int(os.environ.get("LOCAL_RANK", "0"))
    cpus = gpu_cpus(rank)
    if a.bind and cpus:
        os.sched_setaffinity(0, cpus)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    n = int(a.gb * 2 ** 30)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    bw = 5 * n / dt / 1e9
    t = torch.tensor([bw], device="cuda")
    if world > 1:
        lst = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(lst, t)
        if rank == 0:
            v = [float(x) for x in lst]
            print("bind=%s per-GPU H2D GB/s: %s  total %.1f" % (a.bind, " ".join("%.1f" % x for x in v), sum(v)))
        dist.destroy_process_group()
    else:
        print("bind=%s H2D %.1f GB/s (cpus local to GPU %d: %d)" % (a.bind, bw, rank, len(cpus)))
    if rank == 0:
        print("rank0 local cpus:", cpus[:8], "... n=%d of %d; affinity now %d cpus" % (len(cpus), os.cpu_count(), len(os.sched_getaffinity(0))))


if __name__ == "__main__":
    main()
