"""Host-to-host export of the bench workload: zero-copy row gather vs pitched DMA for the input side, same box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
import synth
from sparsespatialsampling_b200.export import KnnTables
from sparsespatialsampling_b200.knn import KnnIndex
N, T = synth.CONFIGS["C2"][0], synth.CONFIGS["C2"][1]
x = synth.cylinder2d_cloud(N, seed=0).cuda()
q = x[pt.randperm(x.size(0), device="cuda")[:58347]] + 1e-4
tables = KnnTables(KnnIndex(x), q, 8)
p = synth.wake_field(x, 0, T, T, 1).cpu().pin_memory()
u = synth.wake_field(x, 0, T, T, 2).cpu().pin_memory()
outs = {}
for mode in (True, False, True, False):
    for chunk in (128, 256):
        def step():
            a = tables.interpolate_host(p, out=outs.get(("p", chunk)), chunk_snapshots=chunk, sync=False, gather=mode)
            b = tables.interpolate_host(u, out=outs.get(("u", chunk)), chunk_snapshots=chunk, sync=False, gather=mode)
            tables.wait_host(); outs[("p", chunk)], outs[("u", chunk)] = a, b
        step(); pt.cuda.synchronize()
        t0 = time.time()
        for _ in range(5): step()
        pt.cuda.synchronize()
        print(f"gather={mode} window {chunk}: {(time.time() - t0) / 5 * 1e3:.2f} ms per step")
