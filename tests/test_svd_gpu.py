"""GPU parity: weighted SVD (s3_svd_row_means / s3_svd_gram / s3_svd_project, compute_svd) against the CPU oracle of
compute_svd (utils.py:302-346)."""
import numpy as np
import pytest
import torch as pt

from oracle import s3_oracle as orc

pytestmark = pytest.mark.gpu

# tolerances: leading singular values rel 1e-4, modes |<u, u_ref>| >= 1 - 1e-4 for well separated modes
RTOL_S = 1e-4
MODE_COS = 1.0 - 1e-4


def _low_rank_field(n_cells, t, n_modes, seed, d=None, noise=1e-3):
    """Snapshots with a few well separated coherent modes + small noise + a large mean, fp32."""
    rng = np.random.default_rng(seed)
    shape = (n_cells,) if d is None else (n_cells, d)
    time = np.linspace(0, 2 * np.pi, t, endpoint=False)
    a = np.zeros(shape + (t,))
    for i in range(n_modes):
        phi = rng.standard_normal(shape)
        a += (2.0 ** -i) * phi[..., None] * np.cos((i + 1) * time + rng.random())
    a += noise * rng.standard_normal(a.shape) + 3.0 + rng.standard_normal(shape)[..., None]
    area = 2.0 ** -rng.integers(8, 16, n_cells).astype(np.float64)
    return a.astype(np.float32), area.astype(np.float32)


@pytest.mark.parametrize("method", ["simt", "tc3", "tc", "tc3-single-cta"])
@pytest.mark.parametrize("m,t", [(3000, 100), (5000, 256), (2500, 301), (1000, 640), (17, 40)])
def test_gram_matches_fp64(cuda, method, m, t):
    from sparsespatialsampling_b200 import svd, _lib
    # default: clusters of two CTAs sharing the B operand by TMA multicast; "-single-cta": one CTA per tile
    _lib.tune(14, 0 if method.endswith("single-cta") else 1)
    method = method.split("-")[0]
    a, area = _low_rank_field(m, t, 5, m + t)
    ref = orc.weighted_gram(a, area)
    ad = pt.from_numpy(a).cuda()
    mean = svd.row_means(ad)
    assert np.allclose(mean.cpu().numpy(), a.mean(axis=1, dtype=np.float64), rtol=1e-6, atol=1e-7)
    g = svd.gram(ad, mean, pt.from_numpy(area).cuda(), 1, method).cpu().numpy()
    assert np.array_equal(g, g.T)
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    tol = {"simt": 2e-5, "tc3": 5e-6, "tc": 3e-3}[method]
    # the mean is rounded to fp32 on both sides; fp32 products/accumulation inside a segment bound the rest
    _lib.tune(14, 1)
    assert (np.abs(g - ref) <= tol * scale + 1e-30).all(), float((np.abs(g - ref) / (scale + 1e-300)).max())


def test_gram_many_rows_segments_and_splits(cuda):
    """K range long enough for several TMEM accumulation segments per CTA (tuning key 10 shrinks the segment)."""
    from sparsespatialsampling_b200 import svd, _lib
    a, area = _low_rank_field(40000, 200, 4, 11)
    ref = orc.weighted_gram(a, area)
    ad = pt.from_numpy(a).cuda()
    mean = svd.row_means(ad)
    vol = pt.from_numpy(area).cuda()
    lib = _lib.load()
    scale = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    try:
        for seg, flush in ((1, 1), (3, 5), (4, 32), (7, 1000), (8, 128)):
            _lib.tune(10, seg)
            _lib.tune(11, flush)
            g = svd.gram(ad, mean, vol, 1, "tc3").cpu().numpy()
            assert (np.abs(g - ref) <= 5e-6 * scale).all(), (seg, flush, float((np.abs(g - ref) / scale).max()))
    finally:
        _lib.tune(10, 8)
        _lib.tune(11, 128)


@pytest.mark.parametrize("d", [None, 2, 3])
@pytest.mark.parametrize("method", ["tc3", "simt"])
def test_compute_svd_against_oracle(cuda, d, method):
    from sparsespatialsampling_b200.svd import compute_svd
    n_cells, t, r = 4000, 120, 6
    a, area = _low_rank_field(n_cells, t, 4, 5, d)
    s_ref, u_ref, v_ref = orc.compute_svd(a, area, rank=r)
    keep = a.copy()
    s, u, v = compute_svd(pt.from_numpy(a), pt.from_numpy(area), rank=r, method=method)
    assert np.array_equal(a, keep), "compute_svd must not modify its input"
    assert s.device.type == "cpu" and tuple(u.shape) == u_ref.shape and tuple(v.shape) == v_ref.shape
    s, u, v = s.numpy(), u.numpy().astype(np.float64), v.numpy().astype(np.float64)
    # 4 coherent modes stand clear of the noise floor
    assert np.allclose(s[:4], s_ref[:4], rtol=RTOL_S)
    w = np.sqrt(area.astype(np.float64))
    w = w[:, None] if d is None else w[:, None, None]
    for i in range(4):
        uw, uw_ref = (u[..., i:i + 1] * w).ravel(), (u_ref[..., i:i + 1].astype(np.float64) * w).ravel()
        assert abs(np.dot(uw, uw_ref)) / (np.linalg.norm(uw) * np.linalg.norm(uw_ref)) >= MODE_COS
        assert abs(np.dot(v[:, i], v_ref[:, i])) >= MODE_COS
        assert abs(np.linalg.norm(uw) - 1.0) < 1e-3           # weighted modes are orthonormal
    # reconstruction of the centred field from the kept triplets matches the oracle's
    flat = (n_cells * (d or 1), t)
    rec = (u.reshape(flat[0], -1) * s[None, :]) @ v.T
    rec_ref = (u_ref.astype(np.float64).reshape(flat[0], -1) * s_ref[None, :]) @ v_ref.T
    assert np.abs(rec - rec_ref).max() <= 1e-3 * np.abs(rec_ref).max()


def test_compute_svd_n_modes_and_device_input(cuda):
    from sparsespatialsampling_b200.svd import compute_svd
    a, area = _low_rank_field(2000, 64, 3, 2)
    s, u, v = compute_svd(pt.from_numpy(a).cuda(), pt.from_numpy(area).cuda(), rank=10, n_modes=2)
    assert u.is_cuda and tuple(u.shape) == (2000, 2) and tuple(s.shape) == (10,) and tuple(v.shape) == (64, 10)
    s2, _, _ = compute_svd(pt.from_numpy(a), pt.from_numpy(area), rank=None)
    assert 3 <= s2.numel() <= 8                                # optimal hard threshold keeps the coherent modes
    with pytest.raises(ValueError):
        compute_svd(pt.from_numpy(a), pt.from_numpy(area[:5]), rank=3)


def test_compute_svd_against_reference_golden(cuda):
    """compute_svd of the reference itself (utils.py:302-346, run in the build container) on a field written by the
    reference's Datawriter: tests/golden/io_golden.npz."""
    import os
    from sparsespatialsampling_b200.svd import compute_svd
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "io_golden.npz"))
    for name in ("p", "U"):
        s, u, v = compute_svd(pt.from_numpy(gold[f"dm_{name}"]), pt.from_numpy(gold["weights"]), rank=4)
        s, u, v = s.numpy(), u.numpy(), v.numpy()
        s_ref = gold[f"svd_{name}_s"]
        assert u.shape == gold[f"svd_{name}_U"].shape
        # compare the modes that are well above the fp32 noise floor of the reference's own SVD
        for i in range(4):
            if s_ref[i] < 1e-3 * s_ref[0]:
                continue
            assert abs(s[i] - s_ref[i]) <= RTOL_S * s_ref[0]
            a, b = u[..., i].ravel().astype(np.float64), gold[f"svd_{name}_U"][..., i].ravel().astype(np.float64)
            assert abs(np.dot(a, b)) / (np.linalg.norm(a) * np.linalg.norm(b)) >= MODE_COS, (name, i)
            assert abs(np.dot(v[:, i], gold[f"svd_{name}_V"][:, i])) >= MODE_COS, (name, i)


def test_write_svd_s_cube_to_file(cuda, tmp_path):
    """End to end: file written by Datawriter -> Dataloader -> device SVD -> modes + XDMF (utils.py:349-413)."""
    import os
    from sparsespatialsampling_b200 import Dataloader, write_svd_s_cube_to_file
    from tests.test_io_format import write_case
    here = os.path.dirname(os.path.abspath(__file__))
    gold = np.load(os.path.join(here, "golden", "io_golden.npz"))
    grid = np.load(os.path.join(here, "golden", "g2d_metric.npz"))
    write_case(str(tmp_path), gold, grid)
    write_svd_s_cube_to_file(["p", "U"], str(tmp_path), "case", new_file=False, n_modes=3, rank=4, t_start=0.2)
    for name in ("p", "U"):
        out = Dataloader(str(tmp_path), f"case_{name}_svd.h5")
        st = out._store()
        assert sorted(st.keys("constant")) == ["V", "cell_area", "mode_1", "mode_2", "mode_3", "s"]
        assert st.read("constant/V").shape == (5, 4) and st.read("constant/s").shape == (4,)    # t >= 0.2: 5 snapshots
        assert st.read("constant/mode_1").shape == ((1535,) if name == "p" else (1535, 2))
        xdmf = open(tmp_path / f"case_{name}_svd.xdmf").read()
        assert 'Attribute Name="mode_3"' in xdmf and 'Attribute Name="cell_area"' in xdmf and "Name=\"V\"" not in xdmf


@pytest.mark.parametrize("m,t,r,vol_div", [(1000, 96, 7, 1), (5003, 301, 40, 1), (3000, 200, 300, 2), (129, 33, 1, 1),
                                           (70000, 1000, 50, 1)])
def test_tensor_core_projection_matches_fp64(cuda, m, t, r, vol_div):
    # s3_svd_project_tc: U = (A - mean) W with the contraction over t on tcgen05 (K-major TF32 operands)
    from sparsespatialsampling_b200 import svd
    g = pt.Generator(device="cuda").manual_seed(m + r)
    a = pt.randn((m, t), device="cuda", generator=g) + 3.0                 # a mean well away from zero
    vol = pt.rand((m // vol_div,), device="cuda", generator=g) + 0.5
    w = pt.randn((t, r), device="cuda", generator=g)
    mean = svd.row_means(a)
    ref = (a.double() - mean.double()[:, None]) @ w.double()
    scale = float(ref.abs().max())
    u3 = svd.project_tc(a, mean, vol, vol_div, w, "tc3")
    assert tuple(u3.shape) == (m, r) and u3.dtype == pt.float32
    assert float((u3.double() - ref).abs().max()) <= 2e-5 * scale           # 3xTF32 ~ fp32 accuracy
    u1 = svd.project_tc(a, mean, vol, vol_div, w, "tc")
    assert float((u1.double() - ref).abs().max()) <= 5e-3 * scale           # single TF32 pass
    us = svd.project(a, mean, w)                                            # fp32 CUDA-core kernel
    assert float((us.double() - ref).abs().max()) <= 2e-5 * scale


def test_small_rank_takes_the_subspace_eigensolver_and_agrees_with_eigh(cuda):
    """compute_svd(rank << T): top-r eigenpairs of the Gram matrix by subspace iteration (svd.top_eigenpairs) instead of
    the full T x T eigh; same factors."""
    from sparsespatialsampling_b200 import svd
    a, area = _low_rank_field(20000, 640, 6, seed=3)
    ad, aread = pt.from_numpy(a).cuda(), pt.from_numpy(area).cuda()
    mean = svd.row_means(ad)
    g = svd.gram(ad, mean, aread.float(), 1, "tc3")
    assert svd.top_eigenpairs(g, 5) is not None                           # converges on the 6 coherent modes ...
    assert svd.top_eigenpairs(g, 40) is None                              # ... and declines the flat noise tail
    outs = {}
    for how in ("auto", "eigh"):
        svd.EIG_METHOD = how
        try:
            outs[how] = svd.compute_svd(ad, aread, rank=5)
        finally:
            svd.EIG_METHOD = "auto"
    (s1, u1, v1), (s2, u2, v2) = outs["auto"], outs["eigh"]
    assert float(((s1 - s2).abs() / s2[0]).max()) < 1e-6
    for i in range(5):
        assert abs(float(v1[:, i] @ v2[:, i])) > 1 - 1e-6
        assert abs(float(u1[:, i] @ u2[:, i])) / float(u1[:, i].norm() * u2[:, i].norm()) > 1 - 1e-5
