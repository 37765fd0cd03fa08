"""Per-step timing and pinned-pool growth of the default (safe) streamed export."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
import synth
from sparsespatialsampling_b200.export import ExportData, KnnTables
from sparsespatialsampling_b200.knn import KnnIndex

N, T = synth.CONFIGS["C2"][0], synth.CONFIGS["C2"][1]
x = synth.cylinder2d_cloud(N, seed=0)
xd = x.cuda()
q = xd[pt.randperm(xd.size(0), device="cuda")[:58347]] + 1e-4
tables = KnnTables(KnnIndex(xd), q, 8)
p = synth.wake_field(xd, 0, T, T, 1).cpu().pin_memory()
u = synth.wake_field(xd, 0, T, T, 2).cpu().pin_memory()


class G:
    pass
g = G()
g.n_dimensions, g.faces, g.vertices, g.levels = 2, None, None, None
g.centers, g.metric, g.size_initial_cell = q, pt.rand(xd.size(0)), 2.2
g.save_path, g.save_name, g.grid_name = "/tmp/s3b200_probe", "c2", "grid"
for async_host in (False, True):
    exp = ExportData(g, write_times=[str(i) for i in range(T)], write_files=False, async_host=async_host)
    exp._tables_centers, exp._initialized_weights, exp._interpolated_metric = tables, True, True
    res_p = res_u = None
    for step in range(8):
        pt.cuda.synchronize()
        t0 = time.time()
        exp.export(x, p, "p")
        t1 = time.time()
        r_p = exp._last_fields.centers
        exp.export(x, u, "U")
        t2 = time.time()
        r_u = exp.interpolated_fields.centers
        t3 = time.time()
        res_p, res_u = r_p, r_u
        print(f"async={async_host} step {step}: export(p) {1e3*(t1-t0):.1f} ms, export(U) {1e3*(t2-t1):.1f} ms, wait {1e3*(t3-t2):.1f} ms, "
              f"total {1e3*(t3-t0):.1f} ms, pool sizes {[len(v) for v in exp._host_pool.values()]}", flush=True)
