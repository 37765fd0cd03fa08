"""
The REFERENCE's own grid generation (sparseSpatialSampling.s_cube.SamplingTree.refine, unmodified, from oracle/_ref) timed
on this host for C1 (t_ref_gridgen) and, with --c2, for the bench workload C2. Writes one JSON
object to stdout; run under gpurun and keep the result as profiles/r2_ref_gridgen_b200.json.
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch as pt
    import synth
    from oracle import reference
    pkg = reference.load()
    if pkg is None:
        print(json.dumps({"unavailable": "oracle/_ref is not installed"}))
        return
    import sparseSpatialSampling.geometry as geo
    from sparseSpatialSampling.s_cube import SamplingTree
    cores = os.cpu_count()
    n_jobs = min(cores, 32)
    pt.set_num_threads(cores)
    out = {"host_cores": cores, "n_jobs": n_jobs, "what": "reference SamplingTree.refine(), unmodified (oracle/_ref)"}
    for name in ["C1"] + (["C2"] if "--c2" in sys.argv else []):
        n = synth.CONFIGS[name][0]
        x = synth.cylinder2d_cloud(n, seed=0)
        m = synth.wake_metric(x)
        geoms = [geo.CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"]),
                 geo.SphereGeometry("cylinder", False, synth.CYL2D["pos"], synth.CYL2D["radius"], refine=True)]
        t0 = time.time()
        tree = SamplingTree(x, m, geoms, n_jobs=n_jobs, uniform_level=5, min_metric=0.75)
        t_setup = time.time() - t0
        tree.refine()
        wall = time.time() - t0
        info = tree.data_final_mesh
        out[name] = {"n_points": int(x.shape[0]), "n_cells": int(info["n_cells"]), "iterations": int(info["iterations"]),
                     "t_total_s": float(info["t_total"]), "t_uniform": float(info["t_uniform"]),
                     "t_adaptive": float(info["t_adaptive"]), "t_geometry": float(info["t_geometry"] or 0.0),
                     "t_renumbering": float(info["t_renumbering"]), "setup_s": t_setup, "wall_s": wall}
        g = os.path.join(ROOT, "tests", "golden", f"config_{name}.npz")
        if os.path.exists(g):
            gold = np.load(g)
            out[name]["same_grid_as_golden"] = bool(int(gold["n_cells"]) == int(info["n_cells"]) and
                                                    int(gold["iterations"]) == int(info["iterations"]))
    print(json.dumps(out))


if __name__ == "__main__":          # the reference spawns a process pool
    main()
