"""
Row-sharded SVD at C5 size under torchrun (one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/svd_sharded.py [--cells 582000] [--snapshots 2000]
Every rank synthesises its own cell window of a rank-12 field + noise, runs compute_svd_sharded twice (second run
timed with CUDA events, max over ranks) and rank 0 prints one JSON line with the stage times.
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
import torch.distributed as dist
from sparsespatialsampling_b200 import svd as S, parallel as P

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=582000)
ap.add_argument("--snapshots", type=int, default=2000)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
pt.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
if world > 1:
    dist.init_process_group("nccl", device_id=pt.device("cuda", pt.cuda.current_device()))
r0, r1 = P.row_window(args.cells, world, rank)
g = pt.Generator(device="cuda").manual_seed(7)
x = (pt.arange(r0, r1, device="cuda", dtype=pt.float32) + 0.5) / args.cells
t = pt.linspace(0, 1, args.snapshots, device="cuda")
a = pt.zeros((r1 - r0, args.snapshots), device="cuda")
for k in range(12):
    a += pt.sin(6.2831853 * (k + 1) * x)[:, None] * pt.cos(6.2831853 * (k + 1) * t + k)[None] / (k + 1)
a += 0.01 * pt.randn(a.shape, device="cuda", generator=pt.Generator(device="cuda").manual_seed(100 + rank))
vol = 0.5 + x


def timed(fn):
    if world > 1:
        dist.barrier()
    pt.cuda.synchronize()
    e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); pt.cuda.synchronize()
    ms = pt.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return out, float(ms)

S.compute_svd_sharded(a, vol, rank=20, n_modes=20)
(s, u, v), total_ms = timed(lambda: S.compute_svd_sharded(a, vol, rank=20, n_modes=20))
mean = S.row_means(a)
gm, gram_ms = timed(lambda: S.gram(a, mean, vol))
_, ar_ms = timed(lambda: P.allreduce_sum(gm))
_, eigh_ms = timed(lambda: pt.linalg.eigh(gm))
if rank == 0:
    print(json.dumps({"n_gpus": world, "cells": args.cells, "snapshots": args.snapshots, "compute_svd_sharded_ms": total_ms,
                      "gram_local_ms": gram_ms, "allreduce_gram_ms": ar_ms, "eigh_ms": eigh_ms,
                      "gram_useful_tflops_aggregate": 2.0 * args.cells * args.snapshots ** 2 / (gram_ms * 1e-3) / 1e12,
                      "s": [float(z) for z in s[:6]]}))
if world > 1:
    dist.destroy_process_group()
