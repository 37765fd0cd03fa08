"""Grouped vs warp-per-cell interpolation kernel on access patterns of increasing regularity (which bound is it?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
from sparsespatialsampling_b200.interpolate import interp_gather, GroupTables
from sparsespatialsampling_b200 import _lib
lib = _lib.load()
for kv in [t for t in (sys.argv[1] if len(sys.argv) > 1 else "").split(",") if t]:
    key, value = kv.split("=")
    _lib.check(lib.s3_set_tuning(int(key), int(value)))
N, T, k = 100000, 1000, 8
dev = "cuda"
data = pt.randn(N, 1, T, device=dev)
w = pt.full((N, k), 1.0 / k, device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); pt.cuda.synchronize()
    return e0.elapsed_time(e1) / 20


def run(name, idx):
    nc = idx.size(0)
    out = pt.empty((nc, 1, T), device=dev)
    g = GroupTables(idx, w[:nc])
    ms_d = timed(lambda: interp_gather(data, idx, w[:nc], out=out))
    ms_g = timed(lambda: g.interpolate(data, out=out))
    u = pt.unique(idx).numel()
    algo = (u + nc) * T * 4 + nc * k * 8
    print(f"{name:40s} rows/cell {g.rows_per_cell:5.2f}  direct {ms_d:.4f} ms {algo / ms_d / 1e6:7.1f} GB/s   grouped {ms_g:.4f} ms {algo / ms_g / 1e6:7.1f} GB/s")


c = pt.arange(N, device=dev, dtype=pt.int64)
ar = pt.arange(k, device=dev)[None, :]
run("same row 8x (pure copy N->N)", c[:, None].repeat(1, k).to(pt.int32).contiguous())
run("8 consecutive rows c..c+7", ((c[:, None] + ar) % N).to(pt.int32).contiguous())
run("rows c + 37*j (scattered, re-used)", ((c[:, None] + 37 * ar) % N).to(pt.int32).contiguous())
run("random rows (no locality)", pt.randint(0, N, (N, k), device=dev, dtype=pt.int32))
run("Nc = N/2, rows 2c..2c+7", ((c[:N // 2, None] * 2 + ar) % N).to(pt.int32).contiguous())
perm = pt.randperm(N, device=dev)
run("8 consecutive rows of a random permutation", perm[(c[:, None] + ar) % N].to(pt.int32).contiguous())
