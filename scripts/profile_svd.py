"""One pass of the SVD kernels at a C5-like size (for ncu): row means, prepare, tensor-core Gram, reduce, projection."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
from sparsespatialsampling_b200 import svd
m, t = int(sys.argv[1]) if len(sys.argv) > 1 else 262144, 2000
if len(sys.argv) > 2:                     # e.g. 11=128,14=0
    from sparsespatialsampling_b200 import _lib
    for kv in sys.argv[2].split(","):
        key, val = kv.split("=")
        _lib.tune(int(key), int(val))
pt.manual_seed(0)
a = pt.randn(m, t, device="cuda")
vol = pt.rand(m, device="cuda") + 0.5
mean = svd.row_means(a)
for _ in range(2):
    g = svd.gram(a, mean, vol, 1, "tc3")
pt.cuda.synchronize()
e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
e0.record(); g = svd.gram(a, mean, vol, 1, "tc3"); e1.record(); pt.cuda.synchronize()
ms = e0.elapsed_time(e1)
tiles, passes = 72, 3
ref = None
if m <= 300000:
    b = (a - mean[:, None]).double() * vol.sqrt().double()[:, None]
    ref = b.T @ b
    err = float(((g - ref).abs() / pt.sqrt(pt.outer(pt.diag(ref), pt.diag(ref)))).max())
    print(f"max rel err vs fp64: {err:.2e}")
print(f"gram tc3 {m}x{t}: {ms:.3f} ms, useful {2.0*m*t*t/ms/1e9:.1f} TFLOP/s, executed tf32 {2.0*m*tiles*128*256*passes/ms/1e9:.1f} TFLOP/s")
u = svd.project(a, mean, pt.randn(t, 16, device="cuda"))
pt.cuda.synchronize()
