"""
Strong scaling of the sharded export on the C4 workload (BASELINE.json config 4: cylinder3D_Re3900-sized synthetic
3-D field, ~10M points, 2000 snapshots, k = 26, snapshot-sharded interpolation at 1/2/4/8 GPUs).

  python scripts/scale_c4.py                                                    (1 GPU, 80 GB of snapshots resident)
  python -m torch.distributed.run --nproc-per-node N ... scripts/scale_c4.py    (rank r holds T/N snapshots)

Rank 0 generates the grid (n_cells_max = 500 000) and the KNN tables, both are broadcast once (NCCL); every rank
generates ITS window of the closed-form field on the device in the pitched layout and exports it through the public
entry point `ExportData(distributed=True).export(coordinates, window, "p", n_snapshots_total=T)` (device tensors in,
device result out, no files). Timing as in bench.py: barrier + synchronize on both sides, CUDA events, max over ranks.
One JSON line from rank 0.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))

import numpy as np
import torch as pt
import torch.distributed as dist

import synth
import sparsespatialsampling_b200 as s3
from sparsespatialsampling_b200.export import ExportData, KnnTables
from sparsespatialsampling_b200.interpolate import alloc_snapshots
from sparsespatialsampling_b200.knn import KnnIndex
from sparsespatialsampling_b200.parallel import broadcast_grid, snapshot_window
from run_config import config, fill_field


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--snapshots", type=int, default=0)
    ap.add_argument("--points", type=int, default=0, help="smaller cloud for quick checks")
    args = ap.parse_args()
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    pt.cuda.set_device(local)
    dev = pt.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        pt.cuda.synchronize()

    def rank_max(v):
        t = pt.tensor([v], dtype=pt.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    if args.points:
        synth.CONFIGS["C4"] = (args.points,) + tuple(synth.CONFIGS["C4"][1:])
    x, geoms, wake, grid_kw = config("C4", 0)
    T = args.snapshots or synth.CONFIGS["C4"][1]
    k = 26
    xd = x.to(dev)
    centers, tables, grid_s, n_unique = None, None, None, 0
    if rank == 0:
        metric = synth.wake_metric(xd, xc=wake["xc"], yc=wake["yc"]).cpu()
        sc = s3.SparseSpatialSampling(x, metric, geoms, "/tmp/s3b200_c4", "c4", **grid_kw)
        sc.execute_grid_generation()
        grid_s = sc.mesh_info["t_total"]
        centers = sc.centers.to(dev)
        tables = KnnTables(KnnIndex(xd), centers, k)
    barrier()
    if world > 1:                                    # NCCL sets its broadcast channels up on first use: not the tables' cost
        KnnTables.share(tables, dev, src=0)
    barrier()
    b0, b1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    b0.record()
    centers = broadcast_grid(centers, 3, dev, src=0)
    tables = KnnTables.share(tables, dev, src=0, n=int(centers.size(0)), k=k)
    b1.record()
    barrier()
    bcast_ms = rank_max(b0.elapsed_time(b1)) if world > 1 else 0.0
    nc = tables.n
    n_unique = int(pt.unique(tables.idx_sorted).numel())

    t0, t1 = snapshot_window(T, world, rank)
    ts = t1 - t0
    data = alloc_snapshots(xd.size(0), 1, ts, device=dev)
    for c0 in range(0, ts, 16):                       # closed-form field of THIS window, generated on the device
        c1 = min(c0 + 16, ts)
        data[:, :, c0:c1] = synth.wake_field(xd, t0 + c0, t0 + c1, T, 1, wake["xc"], wake["yc"])
    class _Grid:
        pass
    g = _Grid()
    g.n_dimensions, g.faces, g.vertices, g.levels = 3, None, None, None
    g.centers, g.metric, g.size_initial_cell = centers, pt.zeros(xd.size(0)), 1.0
    g.save_path, g.save_name, g.grid_name = "/tmp/s3b200_c4", f"c4_rank{rank}", "grid"
    exp = ExportData(g, write_times=[str(i) for i in range(T)], write_files=False, device=dev, distributed=world > 1)
    exp._tables_centers, exp._initialized_weights, exp._interpolated_metric = tables, True, True   # built / shared above

    def step():
        exp.export(x, data, "p", n_snapshots_total=T)
        return exp._last_fields.centers
    for _ in range(max(args.warmup, 3)):
        out = step()
    barrier()
    e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    barrier()
    ms = rank_max(e0.elapsed_time(e1)) / args.steps
    # parity spot check on this rank's window: 64 cells against a direct fp64 evaluation of the same tables
    sel = pt.arange(0, nc, max(1, nc // 64), device=dev)[:64]
    ref = (tables.w64[sel][:, :, None] * data[tables.idx[sel].long(), 0, :].double()).sum(1)
    scale = data[tables.idx[sel].long(), 0, :].abs().amax(1).clamp_min(1e-30)
    ok = bool(((out[sel, 0, :].double() - ref).abs() <= 1e-5 * scale).all())
    flag = pt.tensor([1.0 if ok else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        peak = 6542.7
        p = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(p):
            peak = float(json.load(open(p))["hbm_gbs"])
        b_algo = n_unique * ts * 4 + nc * ts * 4 + nc * k * 8            # per rank
        line = {"workload": "C4: 10M-point 3-D cloud, 2000 snapshots, k = 26, snapshot-sharded", "n_gpus": world,
                "n_points": int(x.size(0)), "n_cells": int(nc), "snapshots_total": T, "snapshots_per_gpu": ts,
                "ms_per_step": ms, "value": nc * T / (ms * 1e-3), "unit": "snapshot-points/s", "scaling": "strong",
                "per_gpu_algorithmic_GBps": b_algo / (ms * 1e-3) / 1e9, "per_gpu_roofline_frac": b_algo / (ms * 1e-3) / 1e9 / peak,
                "tables_broadcast_ms": bcast_ms, "grid_gen_s": grid_s, "unique_source_points": n_unique,
                "parity_within_1e-5_on_all_ranks": bool(flag.item() == 1.0), "steps": args.steps}
        os.dup2(out_fd, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
