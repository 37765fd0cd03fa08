"""cProfile of the C2 grid generation: where the host time of the adaptive loop goes."""
import cProfile, pstats, os, sys, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
import synth
import sparsespatialsampling_b200 as s3
from sparsespatialsampling_b200.s_cube import SamplingTree
x = synth.cylinder2d_cloud(100000, seed=0)
m = synth.wake_metric(x.cuda()).cpu()
geoms = lambda: [s3.geometry.CubeGeometry("domain", True, synth.CYL2D["lower"], synth.CYL2D["upper"]),
                 s3.geometry.SphereGeometry("cylinder", False, synth.CYL2D["pos"], synth.CYL2D["radius"], refine=True)]
SamplingTree(x, m, geoms(), uniform_level=5, min_metric=0.75).refine()
tree = SamplingTree(x, m, geoms(), uniform_level=5, min_metric=0.75)
pr = cProfile.Profile()
pr.enable()
tree.refine()
pr.disable()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28)
print(st.getvalue())
print(tree.data_final_mesh["t_adaptive"], tree.data_final_mesh["iterations"])
