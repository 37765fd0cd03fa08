"""
A/B lab of the interpolation kernel on the bench workload (C2 tables): every launch variant x both layouts, one JSON
line each. Run on the GPU box:  python scripts/interp_lab.py [--variants "2=2;3=1;..."] > gpurun_out/lab.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch as pt

import bench
import synth
import sparsespatialsampling_b200 as s3
from sparsespatialsampling_b200 import _lib
from sparsespatialsampling_b200.export import KnnTables
from sparsespatialsampling_b200.interpolate import alloc_snapshots
from sparsespatialsampling_b200.knn import KnnIndex

DEFAULTS = {1: 8, 2: 0, 3: -1, 4: 0, 5: -1, 6: -1, 7: 0, 8: -1, 9: 1, 12: 0, 13: 0}
VARIANTS = ["", "9=0", "9=0,6=0", "9=0,6=0,2=1", "6=-1,3=0", "2=2", "1=4", "1=16", "4=512", "7=4"]
OLD_VARIANTS = ["", "7=2", "7=4", "2=2", "2=2,7=2", "6=1,3=1", "6=1,3=1,7=2", "8=2", "8=2,7=2", "8=2,7=4", "8=2,4=512",
            "8=2,7=2,4=512", "8=4", "8=4,7=2", "8=4,4=512", "8=4,4=256", "8=4,1=4", "8=4,1=4,4=512", "8=2,1=4",
            "8=2,1=16", "8=4,1=2,4=512"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="")
    ap.add_argument("--snapshots", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--layouts", default="pitched,dense")
    ap.add_argument("--k26", action="store_true", help="3-D style tables (k = 26) on a synthetic 3-D cloud instead of C2")
    args = ap.parse_args()
    dev = pt.device("cuda", 0)
    pt.cuda.set_device(dev)
    T = args.snapshots
    if args.k26:
        x = synth.cylinder3d_cloud(1000000, seed=0)
        xd = x.to(dev)
        g = pt.Generator().manual_seed(1)
        lo, hi = pt.tensor(synth.CYL3D["lower"]), pt.tensor(synth.CYL3D["upper"])
        centers = (lo + pt.rand((200000, 3), generator=g, dtype=pt.float64) * (hi - lo)).to(dev)
        k = 26
    else:
        x = synth.cylinder2d_cloud(bench.N_POINTS, seed=0)
        xd = x.to(dev)
        metric = synth.wake_metric(xd).cpu()
        sc = s3.SparseSpatialSampling(x, metric, bench.geometries(s3.geometry), "/tmp/s3b200_lab", "c2",
                                      uniform_levels=5, min_metric=0.75)
        sc.execute_grid_generation()
        centers = sc.centers.to(dev)
        k = 8
    tables = KnnTables(KnnIndex(xd), centers, k)
    nc = tables.n
    n_unique = int(pt.unique(tables.idx_sorted).numel())
    # rotating buffer sets: at least 500 MB touched between two uses of a byte (126 MB L2), as in bench.py
    per_set = sum((xd.size(0) + nc) * c * T * 4 for c in (1, 2))
    n_sets = max(1, -(-500 * 1024 * 1024 // per_set)) if T < 1000 else 1
    sets = []
    for s_i in range(n_sets):
        fields = {}
        for comps in (1, 2):
            f = synth.wake_field(xd, 0, T, T, comps)
            fields[comps] = {}
            if "dense" in args.layouts:
                fields[comps]["dense"] = (f, pt.empty((nc, comps, T), device=dev))
            if "pitched" in args.layouts:
                fp = alloc_snapshots(xd.size(0), comps, T, device=dev, zero=True)
                fp.copy_(f)
                fields[comps]["pitched"] = (fp, alloc_snapshots(nc, comps, T, device=dev))
        sets.append(fields)
    fields = sets[0]
    b_algo = sum(bench.algorithmic_bytes(n_unique, nc, k, c, T) for c in (1, 2))
    peak, _ = bench.measured_peak()
    variants = [v for v in args.variants.split(";")] if args.variants else VARIANTS
    base = {}
    for var in variants:
        for key, val in DEFAULTS.items():
            _lib.tune(key, val)
        for kv in [t for t in var.split(",") if t]:
            key, val = kv.split("=")
            _lib.tune(int(key), int(val))
        for layout in args.layouts.split(","):
            def step(i=0):
                for comps in (1, 2):
                    d, o = sets[i % n_sets][comps][layout]
                    tables.interpolate(d, pt.float32, out=o)
            for i in range(args.warmup * n_sets):
                step(i)
            pt.cuda.synchronize()
            e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                step(i)
            e1.record()
            pt.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            res = tuple(fields[c][layout][1].clone() for c in (1, 2))
            if not base:
                base["res"] = res
            same = all(pt.equal(a, b) for a, b in zip(res, base["res"]))
            print(json.dumps({"variant": var or "default", "layout": layout, "k": k, "T": T, "n_cells": nc, "buffer_sets": n_sets,
                              "ms_per_step": round(ms, 4), "frac": round(b_algo / (ms * 1e-3) / 1e9 / peak, 4),
                              "bit_identical": same}), flush=True)


if __name__ == "__main__":
    main()
