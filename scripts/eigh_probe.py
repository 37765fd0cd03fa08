import time, torch as pt
for n in (1000, 2000):
    a = pt.randn(n, n, dtype=pt.float64, device="cuda"); g = a @ a.T
    for rep in range(3):
        pt.cuda.synchronize(); t0 = time.time(); w, v = pt.linalg.eigh(g); pt.cuda.synchronize()
        print(f"GPU eigh fp64 n={n} rep {rep}: {time.time()-t0:.3f} s")
    gc = g.cpu()
    t0 = time.time(); w, v = pt.linalg.eigh(gc); print(f"CPU eigh fp64 n={n}: {time.time()-t0:.3f} s ({pt.get_num_threads()} threads)")
    g32 = g.float()
    pt.cuda.synchronize(); t0 = time.time(); w, v = pt.linalg.eigh(g32); pt.cuda.synchronize(); print(f"GPU eigh fp32 n={n}: {time.time()-t0:.3f} s")
