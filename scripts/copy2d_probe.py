"""Rates of the DMA copies of the host-to-host export (C2 geometry) against linear copies, and the e2e step for the
window sizes / transfer modes of KnnTables.interpolate_host.  python scripts/copy2d_probe.py > gpurun_out/copy2d.log"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as pt
import synth
from sparsespatialsampling_b200 import _lib
from sparsespatialsampling_b200.export import KnnTables
from sparsespatialsampling_b200.knn import KnnIndex

lib = _lib.load()
dev = pt.device("cuda", 0)
N, T = synth.CONFIGS["C2"][0], synth.CONFIGS["C2"][1]
NC = 58347
s_a, s_b = pt.cuda.Stream(), pt.cuda.Stream()


def rate(fn, nbytes, reps=4):
    fn(); pt.cuda.synchronize()
    e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    for s in (s_a, s_b):
        pt.cuda.current_stream().wait_stream(s)
    e1.record(); pt.cuda.synchronize()
    return nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


host_in = pt.empty((N * 2, T), dtype=pt.float32, pin_memory=True).normal_()
host_out = pt.empty((NC * 2, T), dtype=pt.float32, pin_memory=True)
for width in (232, 256, 500, 1000):
    pitch = (width * 4 + 127) // 128 * 128 // 4
    d_in = pt.empty((N * 2, pitch), device=dev)
    d_out = pt.empty((NC * 2, pitch), device=dev)

    def h2d():
        s_a.wait_stream(pt.cuda.current_stream())
        _lib.check(lib.s3_copy2d_async(d_in.data_ptr(), pitch * 4, host_in.data_ptr(), T * 4, width * 4, N * 2, 0, s_a.cuda_stream))

    def d2h():
        s_b.wait_stream(pt.cuda.current_stream())
        _lib.check(lib.s3_copy2d_async(host_out.data_ptr(), T * 4, d_out.data_ptr(), pitch * 4, width * 4, NC * 2, 1, s_b.cuda_stream))

    def both():
        h2d(); d2h()
    bi, bo = N * 2 * width * 4, NC * 2 * width * 4
    print(f"2-D copies, {width * 4:5d}-byte rows of a {T * 4}-byte host pitch: h2d {rate(h2d, bi):5.1f} GB/s, d2h {rate(d2h, bo):5.1f} GB/s, "
          f"both {rate(both, bi + bo):5.1f} GB/s", flush=True)
lin_in = pt.empty(N * 2 * T, dtype=pt.float32, pin_memory=True)
lin_d = pt.empty(N * 2 * T, device=dev)
lin_o = pt.empty(NC * 2 * T, dtype=pt.float32, pin_memory=True)
lin_do = pt.empty(NC * 2 * T, device=dev)


def lh2d():
    s_a.wait_stream(pt.cuda.current_stream())
    with pt.cuda.stream(s_a):
        lin_d.copy_(lin_in, non_blocking=True)


def ld2h():
    s_b.wait_stream(pt.cuda.current_stream())
    with pt.cuda.stream(s_b):
        lin_o.copy_(lin_do, non_blocking=True)


def lboth():
    lh2d(); ld2h()
print(f"linear copies of the same sizes: h2d {rate(lh2d, lin_in.numel() * 4):5.1f}, d2h {rate(ld2h, lin_o.numel() * 4):5.1f}, "
      f"both {rate(lboth, (lin_in.numel() + lin_o.numel()) * 4):5.1f} GB/s", flush=True)
del lin_in, lin_d, lin_o, lin_do, host_in, host_out

x = synth.cylinder2d_cloud(N, seed=0).cuda()
q = x[pt.randperm(x.size(0), device="cuda")[:NC]] + 1e-4
tables = KnnTables(KnnIndex(x), q, 8)
p = synth.wake_field(x, 0, T, T, 1).cpu().pin_memory()
u = synth.wake_field(x, 0, T, T, 2).cpu().pin_memory()
bufs = {}


def timeit(fn, reps=5):
    fn(); fn(); pt.cuda.synchronize()
    t0 = time.time()
    for _ in range(reps):
        fn()
    pt.cuda.synchronize()
    return (time.time() - t0) / reps * 1e3


for gather in (False, True):
    for chunk in (128, 256, 334, 500, 1000):
        def step():
            for name, f in (("p", p), ("u", u)):
                bufs[(name, chunk)] = tables.interpolate_host(f, out=bufs.get((name, chunk)), chunk_snapshots=chunk,
                                                              sync=False, gather=gather)
                tables.wait_input()
            tables.wait_host()
        print(f"e2e step, window {chunk:4d}, {'PCIe row gather kernel' if gather else 'pitched DMA copies'}: {timeit(step):6.2f} ms", flush=True)
