"""Host<->device copy rates on this box: contiguous vs pitched windows, one direction vs both (sizes of the C2 step)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
from sparsespatialsampling_b200 import _lib
lib = _lib.load()
N, T = 200000, 1000
h = pt.empty((N, T), dtype=pt.float32).pin_memory(); h.normal_()
h2 = pt.empty((N // 2, T), dtype=pt.float32).pin_memory()
d = pt.empty((N, T), dtype=pt.float32, device="cuda")
d2 = pt.empty((N // 2, T), dtype=pt.float32, device="cuda"); d2.normal_()
s1, s2 = pt.cuda.Stream(), pt.cuda.Stream()

def t(fn, reps=3):
    fn(); pt.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        fn()
    pt.cuda.synchronize()
    return (time.time() - t0) / reps

gb_in, gb_out = N * T * 4 / 1e9, N // 2 * T * 4 / 1e9
dt = t(lambda: d.copy_(h, non_blocking=True)); print(f"H2D contiguous {gb_in / dt:.1f} GB/s")
dt = t(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H contiguous {gb_out / dt:.1f} GB/s")
def both():
    with pt.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with pt.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
dt = t(both); print(f"both directions contiguous: {dt*1e3:.1f} ms, H2D {gb_in / dt:.1f} + D2H {gb_out / dt:.1f} GB/s")
for tc in (32, 64, 128, 256, 512):
    def pitched_in():
        for c in range(0, T, tc):
            w = min(tc, T - c)
            _lib.check(lib.s3_copy2d_async(d.data_ptr(), w * 4, h.data_ptr() + c * 4, T * 4, w * 4, N, 0, s1.cuda_stream))
    dt = t(pitched_in); print(f"H2D pitched windows of {tc} snapshots ({tc*4} B rows): {gb_in / dt:.1f} GB/s")
    def pitched_out():
        for c in range(0, T, tc):
            w = min(tc, T - c)
            _lib.check(lib.s3_copy2d_async(h2.data_ptr() + c * 4, T * 4, d2.data_ptr(), w * 4, w * 4, N // 2, 1, s2.cuda_stream))
    dt = t(pitched_out); print(f"D2H pitched windows of {tc} snapshots: {gb_out / dt:.1f} GB/s")
    def pb():
        pitched_in(); pitched_out()
    dt = t(pb); print(f"  both pitched: {dt*1e3:.1f} ms")
# row-chunked contiguous copies (whole rows) in both directions at once
def rows_both():
    for c in range(0, N, N // 8):
        with pt.cuda.stream(s1): d[c:c + N // 8].copy_(h[c:c + N // 8], non_blocking=True)
        with pt.cuda.stream(s2): h2[c // 2:(c + N // 8) // 2].copy_(d2[c // 2:(c + N // 8) // 2], non_blocking=True)
dt = t(rows_both); print(f"both directions, 8 contiguous row chunks each: {dt*1e3:.1f} ms")
