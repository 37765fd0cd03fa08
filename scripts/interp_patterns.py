"""Where is the interpolation kernel's invariant bound? Same kernel, access patterns of increasing regularity."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
from sparsespatialsampling_b200.interpolate import interp_gather
from sparsespatialsampling_b200 import _lib
lib = _lib.load()
N, T, k = 100000, 1000, 8
dev = "cuda"
data = pt.randn(N, 1, T, device=dev)
w = pt.full((N, k), 1.0 / k, device=dev)

def run(name, idx, out_row=None, nc=None):
    nc = idx.size(0)
    out = pt.empty((nc, 1, T), device=dev)
    for _ in range(3):
        interp_gather(data, idx, w[:nc], out=out, out_row=out_row)
    e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        interp_gather(data, idx, w[:nc], out=out, out_row=out_row)
    e1.record(); pt.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    u = pt.unique(idx).numel()
    algo = (u + nc) * T * 4 + nc * k * 8
    print(f"{name:46s} {ms:.4f} ms  unique {u:6d} cells {nc:6d}  algorithmic {algo / ms / 1e6:7.1f} GB/s  L1-side gathers {nc * k * T * 4 / ms / 1e6:8.1f} GB/s")

c = pt.arange(N, device=dev, dtype=pt.int64)
for unroll in (1, 2):
    _lib.check(lib.s3_set_tuning(5, unroll))
    for regs in (0, 1):
        _lib.check(lib.s3_set_tuning(8, regs))
        print(f"--- unroll {unroll} regs {regs}")
        run("same row 8x (pure copy N->N)", c[:, None].repeat(1, k).to(pt.int32).contiguous())
        run("8 consecutive rows c..c+7", ((c[:, None] + pt.arange(k, device=dev)[None, :]) % N).to(pt.int32).contiguous())
        run("rows c + 37*j (scattered, re-used)", ((c[:, None] + 37 * pt.arange(k, device=dev)[None, :]) % N).to(pt.int32).contiguous())
        run("random rows (no locality)", pt.randint(0, N, (N, k), device=dev, dtype=pt.int32))
        half = ((c[:N // 2, None] * 2 + pt.arange(k, device=dev)[None, :]) % N).to(pt.int32).contiguous()
        run("Nc = N/2, rows 2c..2c+7", half)
        run("Nc = N/2, rows 2c..2c+7, permuted output", half, out_row=pt.randperm(N // 2, device=dev).to(pt.int32))
