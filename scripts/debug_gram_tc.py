"""Debug aid for the tensor-core Gram kernel: random accuracy + structured inputs that expose layout permutations."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch as pt
from sparsespatialsampling_b200 import svd

pt.manual_seed(0)
for (m, t) in [(64, 128), (3000, 100), (5000, 300), (20000, 2000)]:
    a = pt.randn(m, t, device="cuda")
    mean = pt.zeros(m, device="cuda")
    vol = pt.ones(m, device="cuda")
    ref = (a.double().T @ a.double())
    for method in ("simt", "tc", "tc3"):
        g = svd.gram(a, mean, vol, 1, method)
        scale = pt.sqrt(pt.outer(pt.diag(ref), pt.diag(ref)))
        print(m, t, method, "max rel err", float(((g - ref).abs() / scale).max()), "nonzero frac", float((g != 0).float().mean()))

# structured: one row with two non-zeros
m, t = 16, 256
for (m0, t0, t1) in [(0, 0, 1), (0, 0, 5), (1, 0, 33), (3, 2, 64), (5, 7, 130), (9, 31, 200), (15, 100, 255)]:
    a = pt.zeros(m, t, device="cuda")
    a[m0, t0] = 1.0
    a[m0, t1] = 2.0
    g = svd.gram(a, pt.zeros(m, device="cuda"), pt.ones(m, device="cuda"), 1, "tc")
    nz = pt.nonzero(g).tolist()
    print((m0, t0, t1), "expected", sorted([[t0, t0], [t0, t1], [t1, t0], [t1, t1]]), "got", [(i, j, float(g[i, j])) for i, j in nz][:12])

# timing at a C5-like size
def timeit(fn, n=3):
    fn(); pt.cuda.synchronize()
    e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); pt.cuda.synchronize()
    return e0.elapsed_time(e1) / n

m, t = 500000, 2000
a = pt.randn(m, t, device="cuda")
mean = svd.row_means(a)
vol = pt.rand(m, device="cuda") + 0.5
flop = 2.0 * m * t * t
for method in ("tc", "tc3", "simt"):
    ms = timeit(lambda: svd.gram(a, mean, vol, 1, method), 2)
    print(f"gram {method}: {ms:.2f} ms  {flop / ms / 1e9:.1f} TFLOP/s (useful, full square)")
ms = timeit(lambda: svd.row_means(a))
print(f"row means: {ms:.2f} ms {m * t * 4 / ms / 1e6:.0f} GB/s")
b = (a - mean[:, None]) * vol.sqrt()[:, None]
pt.backends.cuda.matmul.allow_tf32 = False
ms = timeit(lambda: b.T @ b, 2)
print(f"torch fp32 matmul (cuBLAS): {ms:.2f} ms {flop / ms / 1e9:.1f} TFLOP/s")
pt.backends.cuda.matmul.allow_tf32 = True
ms = timeit(lambda: b.T @ b, 2)
print(f"torch tf32 matmul (cuBLAS): {ms:.2f} ms {flop / ms / 1e9:.1f} TFLOP/s")
ref = b.double().T @ b.double()
scale = pt.sqrt(pt.outer(pt.diag(ref), pt.diag(ref)))
for method in ("tc", "tc3", "simt"):
    g = svd.gram(a, mean, vol, 1, method)
    print(method, "max rel err vs fp64", float(((g - ref).abs() / scale).max()))
print("cuBLAS tf32 err", float((((b.T @ b).double() - ref).abs() / scale).max()))
vs = pt.randn(t, 16, device="cuda")
ms = timeit(lambda: svd.project(a, mean, vs), 2)
print(f"project r=16: {ms:.2f} ms")
