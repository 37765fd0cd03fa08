"""Rate of pitched host<->device copies (cudaMemcpy2DAsync) as a function of the row width, source pitch 4000 B."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
from sparsespatialsampling_b200 import _lib
lib = _lib.load()
N, T = 100000, 1000
h = pt.empty((N, T), dtype=pt.float32).pin_memory(); h.fill_(1.0)
d = pt.empty((N, T), dtype=pt.float32, device="cuda")
s1 = pt.cuda.Stream()


def t(fn, reps=5):
    fn(); pt.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        fn()
    pt.cuda.synchronize()
    return (time.time() - t0) / reps


for cols in (8, 104, 232, 240, 248, 250, 256, 488, 500, 512, 744, 768, 1000):
    for kind, name in ((0, "H2D"), (1, "D2H")):
        def copy():
            if kind == 0:
                _lib.check(lib.s3_copy2d_async(d.data_ptr(), cols * 4, h.data_ptr(), T * 4, cols * 4, N, 0, s1.cuda_stream))
            else:
                _lib.check(lib.s3_copy2d_async(h.data_ptr(), T * 4, d.data_ptr(), cols * 4, cols * 4, N, 1, s1.cuda_stream))
        dt = t(copy)
        print(f"{name} rows of {cols:4d} snapshots ({cols * 4:4d} B): {N * cols * 4 / dt / 1e9:5.1f} GB/s", end="   ")
    print()
