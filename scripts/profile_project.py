"""Projection U = (A - mean) V / s at C5-like size: tensor-core kernel vs fp32 CUDA-core kernel, time and error."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
from sparsespatialsampling_b200 import svd
m, t = int(sys.argv[1]) if len(sys.argv) > 1 else 262144, 2000
pt.manual_seed(0)
a = pt.randn(m, t, device="cuda") + 1.0
vol = pt.rand(m, device="cuda") + 0.5
mean = svd.row_means(a)


def timed(fn):
    fn(); fn()
    pt.cuda.synchronize()
    e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); pt.cuda.synchronize()
    return out, e0.elapsed_time(e1)


for r in (10, 50, 150, 256, 400):
    w = pt.randn(t, r, device="cuda")
    u3, ms3 = timed(lambda: svd.project_tc(a, mean, vol, 1, w, "tc3"))
    u1, ms1 = timed(lambda: svd.project_tc(a, mean, vol, 1, w, "tc"))
    us, mss = timed(lambda: svd.project(a, mean, w))
    rows = slice(0, 20000)
    ref = (a[rows].double() - mean[rows].double()[:, None]) @ w.double()
    sc = float(ref.abs().max())
    print(f"r={r:4d}: tc3 {ms3:7.3f} ms (err {float((u3[rows].double() - ref).abs().max()) / sc:.1e})  tc {ms1:7.3f} ms "
          f"(err {float((u1[rows].double() - ref).abs().max()) / sc:.1e})  simt {mss:7.3f} ms "
          f"(err {float((us[rows].double() - ref).abs().max()) / sc:.1e})  useful {2.0 * m * t * r / ms3 / 1e9:.0f} TFLOP/s")
