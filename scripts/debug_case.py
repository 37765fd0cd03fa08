import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as pt
from oracle import s3_oracle as orc
from tests.golden.make_golden import case_definitions
import sparsespatialsampling_b200.geometry as geo
from sparsespatialsampling_b200.s_cube import SamplingTree
name = sys.argv[1]
case = case_definitions(geo)[name]
tree = SamplingTree(case["coords"], case["metric"], case["geoms"](geo), **case["kwargs"], sdm_order=1)
tree._selected_log = []
tree.refine()
o = orc.OracleTree(case["coords"].numpy(), case["metric"].numpy(), case["geoms"](geo), **case["kwargs"], sdm_order=1).refine()
print("cells_per_iter gpu", tree.data_final_mesh["cells_per_iter"])
print("cells_per_iter orc", o.n_cells_log)
for it, (a, b) in enumerate(zip(tree._selected_log, o.selected_log)):
    if a != b:
        print("first selection difference at iteration", it, [ (x,y) for x,y in zip(a,b) if x!=y][:10])
        break
else:
    print("selections equal", len(tree._selected_log), len(o.selected_log))
n = min(tree._n_cells, len(o.center))
print("n cells total gpu/orc", tree._n_cells, len(o.center))
cg = tree._center[:n].cpu().numpy(); co = np.stack(o.center[:n])
print("centers equal", np.array_equal(cg, co))
gg = tree._gain[:n].cpu().numpy(); go = np.array(o.gain[:n], dtype=np.float64)
bad = np.nonzero(gg != go)[0]
print("gain mismatches", bad[:20], gg[bad[:5]], go[bad[:5]])
inv_g = (tree._flags[:n].cpu().numpy() & 2) != 0; inv_o = np.array(o.invalid[:n])
bad = np.nonzero(inv_g != inv_o)[0]
print("invalid mismatches", bad[:20], inv_g[bad[:10]], inv_o[bad[:10]])
for b in bad[:4]:
    print(" cell", b, "level", o.level[b], "center", o.center[b], "nodes", o._nodes(b).tolist())
lg, lo = set(tree._leaf_cells), set(o.leaf)
print("leaf only gpu", sorted(lg - lo)[:20], "only orc", sorted(lo - lg)[:20])
