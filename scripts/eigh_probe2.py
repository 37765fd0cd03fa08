"""Cost of the T x T symmetric eigen-decomposition in the forms torch offers (fp64 / fp32, values only, via svd)."""
import torch as pt
pt.manual_seed(0)
for t in (1000, 2000):
    b = pt.randn(4 * t, t, device="cuda", dtype=pt.float64)
    g = b.T @ b


    def timed(fn, reps=3):
        fn(); pt.cuda.synchronize()
        e0, e1 = pt.cuda.Event(enable_timing=True), pt.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record(); pt.cuda.synchronize()
        return out, e0.elapsed_time(e1) / reps

    (lam, vec), ms64 = timed(lambda: pt.linalg.eigh(g))
    g32 = g.float()
    (lam32, vec32), ms32 = timed(lambda: pt.linalg.eigh(g32))
    _, msv = timed(lambda: pt.linalg.eigvalsh(g))
    _, mssvd = timed(lambda: pt.linalg.svd(g, full_matrices=False))
    gc = g.cpu()
    import time
    t0 = time.time(); pt.linalg.eigh(gc); cpu = (time.time() - t0) * 1e3
    err32 = float((lam32.double() - lam).abs().max() / lam.abs().max())
    print(f"T={t}: eigh fp64 {ms64:.1f} ms, fp32 {ms32:.1f} ms (max eigenvalue error {err32:.1e} of the largest), "
          f"eigvalsh fp64 {msv:.1f} ms, svd fp64 {mssvd:.1f} ms, CPU eigh fp64 {cpu:.0f} ms")
