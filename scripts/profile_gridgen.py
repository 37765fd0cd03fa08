"""Per-phase wall-clock breakdown of the adaptive refinement loop (host + device), C2 inputs."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
import synth
import sparsespatialsampling_b200 as s3
from sparsespatialsampling_b200.s_cube import SamplingTree
import logging
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
x = synth.cylinder2d_cloud(n, seed=0)
m = synth.wake_metric(x.cuda()).cpu()
geoms = [s3.geometry.CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"]),
         s3.geometry.SphereGeometry("cylinder", False, synth.CYL2D["pos"], synth.CYL2D["radius"], refine=True)]
acc = {}
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        pt.cuda.synchronize(); t0 = time.perf_counter()
        r = f(*a, **k)
        pt.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
        acc[name + "_n"] = acc.get(name + "_n", 0) + 1
        return r
    setattr(obj, name, g)
for rep in range(2):
    acc.clear()
    tree = SamplingTree(x, m, geoms, uniform_level=5, min_metric=0.75)
    for nm in ["_select", "_refine_cells", "_remove_invalid_cells", "_compute_captured_metric", "_mask"]:
        wrap(tree, nm)
    t0 = time.perf_counter(); tree.refine(); pt.cuda.synchronize(); t1 = time.perf_counter()
    info = tree.data_final_mesh
    print(f"rep {rep}: total {t1-t0:.3f}s t_total {info['t_total']:.3f} uniform {info['t_uniform']:.3f} adaptive {info['t_adaptive']:.3f} "
          f"geometry {info['t_geometry']} renumber {info['t_renumbering']:.3f} iterations {info['iterations']} cells {info['n_cells']}")
    for k_, v in sorted(acc.items()):
        if not k_.endswith("_n"):
            print(f"    {k_:28s} {v*1e3:9.2f} ms over {acc[k_+'_n']} calls = {v*1e3/acc[k_+'_n']:.3f} ms/call")
