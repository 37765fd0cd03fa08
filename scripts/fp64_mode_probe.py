"""Interpolation in the reference's dtype (fp32 snapshots, fp64 weights / accumulation / result) vs the fp32 fast path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
import synth
from sparsespatialsampling_b200.export import KnnTables
from sparsespatialsampling_b200.knn import KnnIndex
dev = pt.device("cuda", 0)
x = synth.cylinder2d_cloud(100000, seed=0)
g = pt.Generator().manual_seed(1)
centers = x[pt.randperm(x.size(0), generator=g)[:58347]] + 1e-4
tables = KnnTables(KnnIndex(x.to(dev)), centers, 8)
T = 1000
p = pt.randn((x.size(0), 1, T), dtype=pt.float32, device=dev)
u = pt.randn((x.size(0), 2, T), dtype=pt.float32, device=dev)


def timed(dtype):
    outs = [pt.empty((tables.n, c, T), dtype=dtype, device=dev) for c in (1, 2)]
    for _ in range(3):
        tables.interpolate(p, dtype, out=outs[0]); tables.interpolate(u, dtype, out=outs[1])
    e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tables.interpolate(p, dtype, out=outs[0]); tables.interpolate(u, dtype, out=outs[1])
    e1.record(); pt.cuda.synchronize()
    return e0.elapsed_time(e1) / 20, outs


ms32, o32 = timed(pt.float32)
ms64, o64 = timed(pt.float64)
uniq = int(pt.unique(tables.idx_sorted).numel())
b32 = (uniq + tables.n) * 3 * T * 4 + 2 * tables.n * 8 * 8
b64 = uniq * 3 * T * 4 + tables.n * 3 * T * 8 + 2 * tables.n * 8 * 12
print(f"fp32 path {ms32:.3f} ms ({b32 / ms32 / 1e6:.0f} GB/s algorithmic), fp64-result path {ms64:.3f} ms "
      f"({b64 / ms64 / 1e6:.0f} GB/s algorithmic), max |fp32 - fp64| {float((o32[1].double() - o64[1]).abs().max()):.2e}")
