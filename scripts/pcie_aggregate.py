"""
Aggregate host<->device bandwidth of the box with N ranks copying at the same time (torchrun): contiguous pinned
buffers of 512 MB, H2D alone, D2H alone, both directions. Rank 0 prints the per-rank mean and the total.
  python -m torch.distributed.run --nproc-per-node N scripts/pcie_aggregate.py
"""
import os
import time

import torch as pt
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
pt.cuda.set_device(local)
dev = pt.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 128 * 1024 * 1024
h_in = pt.empty(n, dtype=pt.float32, pin_memory=True).normal_()
h_out = pt.empty(n, dtype=pt.float32, pin_memory=True)
d_in = pt.empty(n, dtype=pt.float32, device=dev)
d_out = pt.empty(n, dtype=pt.float32, device=dev).normal_()
s1, s2 = pt.cuda.Stream(), pt.cuda.Stream()


def sync():
    pt.cuda.synchronize()
    if world > 1:
        dist.barrier()


def run(kind, reps=4):
    def once():
        if kind in ("h2d", "both"):
            with pt.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if kind in ("d2h", "both"):
            with pt.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    once()
    sync()
    t0 = time.time()
    for _ in range(reps):
        once()
    pt.cuda.synchronize()
    dt = (time.time() - t0) / reps
    gb = n * 4 / 1e9 * (2 if kind == "both" else 1)
    rate = pt.tensor([gb / dt], device=dev)
    if world > 1:
        dist.all_reduce(rate)
    sync()
    return rate.item()


for kind in ("h2d", "d2h", "both"):
    total = run(kind)
    if rank == 0:
        print(f"{world} ranks, {kind}: {total / world:.1f} GB/s per rank, {total:.1f} GB/s in total", flush=True)
if world > 1:
    dist.destroy_process_group()
