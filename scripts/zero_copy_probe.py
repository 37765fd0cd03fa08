"""PCIe read rate of an SM kernel gathering rows from pinned host memory (s3_gather_rows) vs the DMA engine."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
from sparsespatialsampling_b200 import _lib
lib = _lib.load()
N, T = 100000, 1000
h = pt.empty((N, T), dtype=pt.float32).pin_memory(); h.normal_()
d = pt.empty((N, 256), dtype=pt.float32, device="cuda")
def t(fn, reps=5):
    fn(); pt.cuda.synchronize(); t0 = time.time()
    for _ in range(reps): fn()
    pt.cuda.synchronize(); return (time.time() - t0) / reps
for frac in (1.0, 0.88, 0.3):
    rows = pt.randperm(N)[:int(N * frac)].sort().values.to(pt.int32).cuda()
    for w in (128, 256):
        for ctas in (74, 148, 296, 592):
            def go():
                _lib.check(lib.s3_gather_rows(h.data_ptr(), T, _lib.ptr(rows), rows.numel(), w, _lib.ptr(d), 256, ctas, _lib.stream_ptr()))
            dt = t(go)
            print(f"rows {frac:.2f} window {w} ctas {ctas}: {rows.numel() * w * 4 / dt / 1e9:.1f} GB/s ({dt*1e3:.2f} ms)")
# check correctness
rows = pt.arange(0, N, 3, dtype=pt.int32).cuda()
_lib.check(lib.s3_gather_rows(h.data_ptr() + 4 * 256, T, _lib.ptr(rows), rows.numel(), 256, _lib.ptr(d), 256, 0, _lib.stream_ptr()))
pt.cuda.synchronize()
assert pt.equal(d[:rows.numel()].cpu(), h[::3, 256:512]), "gather from pinned host memory differs"
print("gather from pinned host memory: values ok")
