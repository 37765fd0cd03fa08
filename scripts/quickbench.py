import torch as pt, numpy as np, time, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparsespatialsampling_b200.knn import KnnIndex
from sparsespatialsampling_b200.interpolate import interp_gather
pt.manual_seed(0)
for (N, dim, k, Nc, T) in [(100000, 2, 8, 30000, 1000), (1000000, 2, 8, 100000, 2000), (10000000, 3, 26, 500000, 128)]:
    X = pt.rand(N, dim, dtype=pt.float64, device="cuda")
    t0 = time.time(); ix = KnnIndex(X); pt.cuda.synchronize(); t1 = time.time()
    Q = pt.rand(Nc, dim, dtype=pt.float64, device="cuda")
    ix.tables(Q, k); pt.cuda.synchronize()
    t2 = time.time(); idx, w32, w64 = ix.tables(Q, k); pt.cuda.synchronize(); t3 = time.time()
    print(f"N={N} dim={dim}: build {t1-t0:.4f}s  tables({Nc}) {t3-t2:.4f}s")
    # morton-ish order of queries: sort by idx[:,0]
    data = pt.randn(N, 1, T, device="cuda")
    out = interp_gather(data, idx, w32)
    order = pt.argsort(idx[:, 0])
    idx_s, w_s = idx[order].contiguous(), w32[order].contiguous()
    orow = order.to(pt.int32)
    for name, (i_, w_, r_) in {"unsorted": (idx, w32, None), "sorted": (idx_s, w_s, orow)}.items():
        for _ in range(3): interp_gather(data, i_, w_, out=out, out_row=r_)
        e0, e1 = pt.cuda.Event(True), pt.cuda.Event(True)
        e0.record()
        for _ in range(10): interp_gather(data, i_, w_, out=out, out_row=r_)
        e1.record(); pt.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        U = pt.unique(idx).numel()
        algo = (U + Nc) * T * 4 + Nc * k * 8
        print(f"   interp {name}: {ms:.3f} ms  U={U} algo={algo/1e6:.1f} MB -> {algo/ms/1e6:.1f} GB/s; naive {(Nc*k+Nc)*T*4/ms/1e6:.1f} GB/s")
    del data, out
