"""
Seeded synthetic CFD-like inputs for the five BASELINE.json configurations (no real CFD data is available offline).

Used by bench.py, the tests and tests/golden/make_golden.py, so that the GPU path, the CPU oracle and the reference
all see the same point clouds and snapshot fields. Coordinates are fp64 without duplicate points, snapshot fields
fp32 ``[N, D, T]`` (the reference's layout, sparseSpatialSampling/export.py:135-141). All fields are closed forms of
(x, t), so large cases can be generated shard by shard directly on the device.
"""
import math

import numpy as np
import torch as pt

CYL2D = dict(lower=[0.0, 0.0], upper=[2.2, 0.41], pos=[0.2, 0.2], radius=0.05)
CYL3D = dict(lower=[0.0, 0.0, 0.0], upper=[2.4, 2.0, math.pi * 0.1], pos=[0.8, 1.0], radius=0.05)
AIRFOIL2D = dict(lower=[-1.0, -1.0], upper=[3.0, 1.0])

CONFIGS = {
    # name: (n_points, n_snapshots, dim)
    "C1": (20000, 100, 2),
    "C2": (100000, 1000, 2),
    "C3": (1000000, 2000, 2),
    "C4": (10000000, 2000, 3),
    "C5": (10000000, 2000, 3),
}


def cylinder2d_cloud(n: int, seed: int = 0) -> pt.Tensor:
    """Uniform-random points in [0,2.2]x[0,0.41] minus the disc r=0.05 at (0.2,0.2); fp64 [~n, 2]."""
    g = np.random.default_rng(seed)
    lo, hi = np.array(CYL2D["lower"]), np.array(CYL2D["upper"])
    x = lo + g.random((n, 2)) * (hi - lo)
    keep = np.hypot(x[:, 0] - CYL2D["pos"][0], x[:, 1] - CYL2D["pos"][1]) > CYL2D["radius"]
    return pt.from_numpy(np.ascontiguousarray(x[keep]))


def cylinder3d_cloud(n: int, seed: int = 0) -> pt.Tensor:
    """Uniform-random points in the cylinder3D box minus the cylinder d=0.1 along z at (0.8, 1.0); fp64 [~n, 3]."""
    g = np.random.default_rng(seed)
    lo, hi = np.array(CYL3D["lower"]), np.array(CYL3D["upper"])
    x = lo + g.random((n, 3)) * (hi - lo)
    keep = np.hypot(x[:, 0] - CYL3D["pos"][0], x[:, 1] - CYL3D["pos"][1]) > CYL3D["radius"]
    return pt.from_numpy(np.ascontiguousarray(x[keep]))


def airfoil2d_cloud(n: int, seed: int = 0) -> pt.Tensor:
    """Points in [-1,3]x[-1,1] clustered towards a wedge 'airfoil' spanning x in [0,1] (half of them)."""
    g = np.random.default_rng(seed)
    lo, hi = np.array(AIRFOIL2D["lower"]), np.array(AIRFOIL2D["upper"])
    far = lo + g.random((n // 2, 2)) * (hi - lo)
    near = np.stack([g.random(n - n // 2) * 1.6 - 0.3, g.standard_normal(n - n // 2) * 0.15], 1)
    near = np.clip(near, lo + 1e-9, hi - 1e-9)
    x = np.concatenate([far, near], 0)
    # drop points inside the wedge (0,0)-(1,0.06)-(1,-0.06)
    inside = (x[:, 0] > 0) & (x[:, 0] < 1) & (np.abs(x[:, 1]) < 0.06 * x[:, 0])
    return pt.from_numpy(np.ascontiguousarray(x[~inside]))


def _amplitude(x: pt.Tensor, y: pt.Tensor, xc: float, yc: float) -> pt.Tensor:
    """Wake envelope behind a body at (xc, yc): grows over ~0.2 downstream, spreads linearly."""
    s = (x - xc).clamp(min=0.0)
    sigma = 0.05 + 0.12 * s
    return pt.exp(-((y - yc) / sigma) ** 2) * (1.0 - pt.exp(-s / 0.2))


def wake_field(coords: pt.Tensor, t0: int, t1: int, n_total: int, components: int = 1, xc: float = 0.2,
               yc: float = 0.2, dtype=pt.float32) -> pt.Tensor:
    """
    Travelling-wave wake, snapshots t0..t1-1 of n_total: ``[N, components, t1 - t0]`` on the device of ``coords``.
    components=1: pressure-like scalar; 2/3: velocity-like vector.
    """
    x, y = coords[:, 0:1], coords[:, 1:2]
    amp = _amplitude(x, y, xc, yc)
    t = pt.arange(t0, t1, device=coords.device, dtype=coords.dtype).unsqueeze(0) / max(n_total, 1)
    phase = 2.0 * math.pi * (x / 0.35 - 8.0 * t)
    side = pt.tanh((y - yc) / 0.05)
    if coords.size(1) == 3:
        amp = amp * (1.0 + 0.3 * pt.sin(2.0 * math.pi * coords[:, 2:3] / (math.pi * 0.1)))
    p = amp * pt.sin(phase) + 0.2 * amp * side * pt.sin(2.0 * phase)
    if components == 1:
        return p.unsqueeze(1).to(dtype)
    comps = [1.0 - 0.5 * amp * pt.cos(phase), 0.5 * amp * side * pt.sin(phase)]
    if components == 3:
        comps.append(0.1 * amp * pt.sin(phase + 1.0))
    return pt.stack(comps[:components], dim=1).to(dtype)


def wake_metric(coords: pt.Tensor, n_snapshots: int = 64, xc: float = 0.2, yc: float = 0.2) -> pt.Tensor:
    """std over time of the scalar wake field, fp64 [N] (evaluated from ``n_snapshots`` snapshots)."""
    out = pt.empty(coords.size(0), dtype=pt.float64, device=coords.device)
    step = 1 << 18
    for s in range(0, coords.size(0), step):
        f = wake_field(coords[s:s + step], 0, n_snapshots, n_snapshots, 1, xc, yc, dtype=pt.float64)
        out[s:s + step] = f.squeeze(1).std(dim=1)
    return out
