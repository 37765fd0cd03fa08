"""
STL mask kernel: tiled node-per-thread kernel (csrc/stl.cuh) against the per-cell-thread path (in_stl over all triangles)
on a C5-like batch of cells, for two surface resolutions. One JSON line per case.
  python scripts/stl_bench.py            (ncu: -k regex:stl_inside_kernel)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch as pt

from sparsespatialsampling_b200 import _lib
from sparsespatialsampling_b200.geometry import CubeGeometry, GeometrySTL3D
from sparsespatialsampling_b200.geometry.device import GeometryTable
from tests.stl_util import icosphere_triangles, write_binary_stl

dev = pt.device("cuda")
lib = _lib.load()
rng = np.random.default_rng(0)
n_cells = 80000                                        # one adaptive iteration of C4 / C5 creates ~80k cells
lo, hi = np.array([0.0, 0.0, 0.0]), np.array([2.4, 2.0, 0.3141592653589793])
body_c = np.array([0.8, 1.0, hi[2] / 2])
# half of the cells around the body (where the refinement happens), half anywhere
near = body_c + rng.standard_normal((n_cells // 2, 3)) * 0.12
cen = np.concatenate([near, lo + rng.random((n_cells - n_cells // 2, 3)) * (hi - lo)])
center = pt.from_numpy(np.clip(cen, lo, hi)).to(dev)
level = pt.from_numpy(rng.integers(5, 9, n_cells).astype(np.int32)).to(dev)
for sub in (4, 6):
    path = f"/tmp/s3b200_stl_bench_{sub}.stl"
    write_binary_stl(path, icosphere_triangles(sub, 0.12, tuple(body_c)))
    geoms = [CubeGeometry("domain", True, lo.tolist(), hi.tolist()), GeometrySTL3D("body", False, path, refine=True)]
    tab = GeometryTable(geoms, dev)
    n_tri = int(tab._stl_meta[1, 1])
    res = {}
    for label, stl_geoms, meta in (("tiled", tab.stl_geoms, tab.stl_meta), ("per_thread", 0, None)):
        inv = pt.empty(n_cells, dtype=pt.uint8, device=dev)

        def run():
            _lib.check(lib.s3_cells_mask(_lib.ptr(center), _lib.ptr(level), None, 0, n_cells, 3, 2.4, _lib.ptr(tab.hdr),
                                         _lib.ptr(tab.par), tab.n, -1, 0, 0, _lib.ptr(inv), None, None, stl_geoms, meta,
                                         _lib.stream_ptr()))
        run()
        pt.cuda.synchronize()
        reps = 5 if label == "tiled" else 1
        e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record()
        pt.cuda.synchronize()
        res[label] = (e0.elapsed_time(e1) / reps, inv.cpu())
    assert pt.equal(res["tiled"][1], res["per_thread"][1])
    pairs = n_cells * 8 * n_tri
    print(json.dumps({"triangles": n_tri, "cells": n_cells, "nodes": n_cells * 8, "invalid_cells": int(res["tiled"][1].sum()),
                      "tiled_ms": res["tiled"][0], "per_thread_ms": res["per_thread"][0],
                      "speedup": res["per_thread"][0] / res["tiled"][0],
                      "tiled_ns_per_1e6_triangle_node_pairs": res["tiled"][0] * 1e6 / (pairs / 1e6),
                      "per_thread_ns_per_1e6_triangle_node_pairs": res["per_thread"][0] * 1e6 / (pairs / 1e6)}), flush=True)
