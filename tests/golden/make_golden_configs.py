"""
Golden digests of the REFERENCE's grid generation at the BASELINE.json configuration sizes that its CPU implementation
can finish in this container (C1: ~20k points, C2: ~100k points -- the bench workload). Run once by hand:

    python tests/golden/make_golden_configs.py C1 [C2]

Stores, per configuration, SHA-256 digests of the reference's centers / levels / faces / vertices plus the small
per-iteration logs (tests/golden/config_<name>.npz); the GPU tests compare digests of their own outputs.
"""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def config(name, geo):
    import synth
    n = synth.CONFIGS[name][0]
    x = synth.cylinder2d_cloud(n, seed=0)
    geoms = [geo.CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"]),
             geo.SphereGeometry("cylinder", False, synth.CYL2D["pos"], synth.CYL2D["radius"], refine=True)]
    return x, synth.wake_metric(x), geoms, dict(uniform_level=5, min_metric=0.75)


def main():
    sys.path.insert(0, HERE)
    from make_golden import write_stubs
    stubs = write_stubs()
    sys.path[:0] = [stubs, REF, ROOT]
    os.environ["PYTHONPATH"] = os.pathsep.join([stubs, REF, ROOT, os.environ.get("PYTHONPATH", "")])
    import sparseSpatialSampling.geometry as geo_ref
    from sparseSpatialSampling.s_cube import SamplingTree
    for name in sys.argv[1:]:
        x, m, geoms, kw = config(name, geo_ref)
        t0 = time.time()
        tree = SamplingTree(x, m, geoms, n_jobs=8, **kw)
        tree.refine()
        wall = time.time() - t0
        info = tree.data_final_mesh
        leaf = list(tree._leaf_cells)
        assert leaf == sorted(leaf)
        out = dict(
            n_points=np.int64(x.shape[0]), n_cells=np.int64(info["n_cells"]), iterations=np.int64(info["iterations"]),
            cells_per_iter=np.asarray(info["cells_per_iter"], dtype=np.int64),
            metric_per_iter=np.asarray(info["metric_per_iter"], dtype=np.float64),
            leaf_index_sha=np.array(digest(np.asarray(leaf, dtype=np.int64))),
            centers_sha=np.array(digest(tree.all_centers.numpy())), levels_sha=np.array(digest(tree.all_levels.numpy())),
            faces_sha=np.array(digest(tree.face_ids.numpy())), vertices_sha=np.array(digest(tree.all_nodes.numpy())),
            n_vertices=np.int64(tree.all_nodes.shape[0]), faces_dtype=np.array(str(tree.face_ids.numpy().dtype)),
            reference_t_total=np.float64(info["t_total"]), reference_wall_s=np.float64(wall),
            reference_times=np.asarray([info["t_uniform"], info["t_adaptive"], info["t_geometry"] or 0.0,
                                        info["t_renumbering"]]),
        )
        np.savez_compressed(os.path.join(HERE, f"config_{name}.npz"), **out)
        print(f"wrote config_{name}.npz: {int(out['n_cells'])} cells, {int(out['iterations'])} iterations, "
              f"reference t_total {info['t_total']:.1f} s (wall {wall:.1f} s, n_jobs=8)")


if __name__ == "__main__":
    main()
