"""GPU edge cases of the C-ABI entry points: empty and minimal inputs, k equal to the number of points, ragged row
lengths, invalid arguments (status codes + s3_last_error, nothing throws across the boundary)."""
import ctypes

import numpy as np
import pytest
import torch as pt

from oracle import s3_oracle as orc

pytestmark = pytest.mark.gpu


def test_empty_and_minimal_interpolation(cuda):
    from sparsespatialsampling_b200.interpolate import interp_gather
    data = pt.randn(10, 1, 5, device="cuda")
    out = interp_gather(data, pt.zeros((0, 8), dtype=pt.int32, device="cuda"), pt.zeros((0, 8), device="cuda"))
    assert tuple(out.shape) == (0, 1, 5)
    one = interp_gather(data, pt.tensor([[3]], dtype=pt.int32, device="cuda"), pt.tensor([[1.0]], device="cuda"))
    assert pt.equal(one[0], data[3])
    # a single snapshot (row length 1) and a row length that is no multiple of the vector width
    for t in (1, 3, 130):
        d = pt.randn(50, 2, t, device="cuda")
        idx = pt.randint(0, 50, (20, 8), device="cuda", dtype=pt.int32)
        w = pt.rand(20, 8, device="cuda")
        got = interp_gather(d, idx, w)
        want = (w[:, :, None, None].double() * d[idx.long()].double()).sum(1)
        assert pt.allclose(got.double(), want, rtol=1e-5, atol=1e-6)


def test_knn_with_k_equal_to_the_number_of_points_and_no_queries(cuda):
    from sparsespatialsampling_b200.knn import KnnIndex
    rng = np.random.default_rng(0)
    for dim, n in ((2, 8), (3, 26), (2, 1)):
        x = rng.random((n, dim))
        q = rng.random((5, dim))
        index = KnnIndex(pt.from_numpy(x).cuda())
        dist, idx = index.kneighbors(pt.from_numpy(q).cuda(), n)
        d_ref, i_ref = orc.knn_search_numpy(x, q, n)
        assert np.array_equal(idx.cpu().numpy(), i_ref) and np.array_equal(dist.cpu().numpy(), d_ref)
        dist0, idx0 = index.kneighbors(pt.zeros((0, dim), dtype=pt.float64, device="cuda"), min(n, 3))
        assert tuple(idx0.shape) == (0, min(n, 3))


def test_invalid_arguments_return_status_codes(cuda):
    from sparsespatialsampling_b200 import _lib
    lib = _lib.load()
    # NULL pointers
    assert lib.s3_interp_gather(None, 0, 1, 1, None, None, 1, 1, None, None, 0, None) != 0
    assert b"NULL" in lib.s3_last_error()
    # k larger than the number of points
    x = pt.rand(5, 2, dtype=pt.float64, device="cuda")
    h = ctypes.c_void_p()
    _lib.check(lib.s3_knn_build(_lib.ptr(x), 5, 2, None, _lib.stream_ptr(), ctypes.byref(h)))
    idx = pt.empty((1, 9), dtype=pt.int64, device="cuda")
    dist = pt.empty((1, 9), dtype=pt.float64, device="cuda")
    rc = lib.s3_knn_query(h, _lib.ptr(x[:1].contiguous()), 1, 9, _lib.ptr(idx), _lib.ptr(dist), _lib.stream_ptr())
    assert rc != 0 and lib.s3_last_error()
    _lib.check(lib.s3_knn_free(h))
    # unknown tuning key, selection of more cells than exist, bad SVD method
    assert lib.s3x_tune(999, 1) != 0
    g = pt.rand(4, dtype=pt.float64, device="cuda")
    f = pt.ones(4, dtype=pt.uint8, device="cuda")
    o = pt.empty(8, dtype=pt.int64, device="cuda")
    assert lib.s3_select_topk(_lib.ptr(g), _lib.ptr(f), 4, 8, _lib.ptr(o), _lib.stream_ptr()) != 0
    a = pt.rand(4, 4, device="cuda")
    m = pt.zeros(4, device="cuda")
    gm = pt.empty((4, 4), dtype=pt.float64, device="cuda")
    assert lib.s3_svd_gram(_lib.ptr(a), _lib.ptr(m), _lib.ptr(m), 1, 4, 4, 7, _lib.ptr(gm), _lib.stream_ptr()) != 0
    with pytest.raises(_lib.S3Error):
        _lib.tune(999, 1)


def test_svd_of_tiny_and_single_snapshot_matrices(cuda):
    from sparsespatialsampling_b200 import svd
    for m, t in ((1, 1), (3, 2), (40, 1), (33, 5)):
        a = pt.randn(m, t, device="cuda")
        vol = pt.rand(m, device="cuda") + 0.1
        mean = svd.row_means(a)
        ref = ((a - mean[:, None]).double() * vol.sqrt().double()[:, None])
        ref = ref.T @ ref
        for method in ("tc3", "tc", "simt"):
            g = svd.gram(a, mean, vol, 1, method)
            # single-pass TF32 ("tc") rounds the inputs to 11 bits: 3e-3 of the largest entry; the others 1e-5
            atol = (3e-3 if method == "tc" else 1e-5) * float(ref.abs().max() + 1e-30) + 1e-12
            assert pt.allclose(g, ref, rtol=1e-3, atol=atol), (m, t, method)


def test_export_accepts_two_dimensional_scalar_field(cuda, tmp_path):
    # export.py:186-190: a [N, T] field is reshaped to [N, 1, T] with a warning
    from sparsespatialsampling_b200.export import ExportData
    rng = np.random.default_rng(1)

    class _Grid:
        pass
    g = _Grid()
    g.n_dimensions, g.faces, g.vertices, g.levels = 2, None, None, None
    g.centers, g.metric, g.size_initial_cell = pt.from_numpy(rng.random((300, 2))), pt.from_numpy(rng.random(1000)), 1.0
    g.save_path, g.save_name, g.grid_name = str(tmp_path), "c", "grid"
    x = pt.from_numpy(rng.random((1000, 2)))
    data = pt.from_numpy(rng.standard_normal((1000, 6)).astype(np.float32))
    exp = ExportData(g, write_times=[str(i) for i in range(6)], write_files=False)
    exp.export(x, data.cuda(), "p")
    assert tuple(exp.interpolated_fields.centers.shape) == (300, 1, 6)
    with pytest.raises(ValueError):
        exp.export(x, data[:, 0].cuda(), "p")
    exp2 = ExportData(g, write_files=False)
    with pytest.raises(ValueError):
        exp2.export(x, data.cuda(), "p")           # no write_times


@pytest.mark.parametrize("m,t,d", [(1, 1, 0), (5, 1, 0), (5, 3, 0), (1, 7, 0), (40, 2, 2), (300, 5, 3), (2, 300, 0),
                                   (1000, 17, 0)])
@pytest.mark.parametrize("method", ["tc3", "tc", "simt"])
def test_svd_degenerate_shapes(cuda, m, t, d, method):
    # one snapshot, one cell, fewer cells than snapshots, snapshot counts that are no multiple of the TMA / UMMA tiles
    import numpy as np
    from oracle import s3_oracle as orc
    from sparsespatialsampling_b200 import compute_svd
    rng = np.random.default_rng(m * 31 + t)
    a = rng.standard_normal((m, t) if d == 0 else (m, d, t)).astype(np.float32)
    vol = rng.random(m) + 0.5
    s, u, v = compute_svd(pt.from_numpy(a).cuda(), pt.from_numpy(vol), rank=3, method=method)
    r = s.numel()
    assert r == min(3, t, m * max(d, 1))
    s_ref, _, _ = orc.compute_svd(a, vol, r)
    tol = (5e-3 if method == "tc" else 2e-4) * max(float(s_ref[0]), 1e-12)
    assert np.allclose(s.cpu().numpy(), s_ref[:r], atol=tol)
    assert tuple(u.shape) == ((m, r) if d == 0 else (m, d, r)) and tuple(v.shape) == (t, r)
    assert bool(pt.isfinite(u).all() and pt.isfinite(v).all() and pt.isfinite(s).all())
