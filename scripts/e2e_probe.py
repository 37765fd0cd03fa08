"""Where the time of the host-to-host export goes (C2 step: p + U, 1.19 GB in, 0.70 GB out): pinned allocation, the
wait for the input reads, window size, buffer re-use."""
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


def timeit(fn, reps=5):
    fn(); pt.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        fn()
    pt.cuda.synchronize()
    return (time.time() - t0) / reps * 1e3


def alloc():
    a = pt.empty((58347, 1, T), dtype=pt.float32, pin_memory=True)
    b = pt.empty((58347, 2, T), dtype=pt.float32, pin_memory=True)
    return a, b
print(f"fresh pinned result tensors (dropped at once): {timeit(alloc):.2f} ms per step")
held = []
def alloc_hold():
    held.append(alloc())
    if len(held) > 1:
        held.pop(0)
print(f"fresh pinned result tensors (previous step's still alive): {timeit(alloc_hold):.2f} ms per step")

bufs = {}
for chunk in (128, 256, 512):
    for mode in ("reuse+async", "reuse+wait_input", "fresh+async", "fresh+wait_input"):
        def step():
            outs = []
            for name, f in (("p", p), ("u", u)):
                out = bufs.get((name, chunk)) if mode.startswith("reuse") else None
                r = tables.interpolate_host(f, out=out, chunk_snapshots=chunk, sync=False, gather=False)
                if mode.startswith("reuse"):
                    bufs[(name, chunk)] = r
                if mode.endswith("wait_input"):
                    tables.wait_input()
                outs.append(r)
            tables.wait_host()
            return outs
        print(f"window {chunk:4d} {mode:18s}: {timeit(step):.2f} ms per step", flush=True)
