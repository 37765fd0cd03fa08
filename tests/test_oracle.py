"""CPU tests: pin the oracle against scikit-learn (the reference's third-party KNN) and torch."""
import numpy as np
import pytest
import torch as pt

from oracle import s3_oracle as orc


@pytest.mark.parametrize("dim,k", [(2, 8), (3, 26), (2, 5), (3, 11)])
def test_knn_and_idw_match_sklearn(dim, k):
    # reference call sites: s_cube.py:161-163 (fit), :224 (predict)
    from sklearn.neighbors import KNeighborsRegressor
    rng = np.random.default_rng(dim * 100 + k)
    X = rng.random((20000, dim))
    y = rng.random(20000)
    Q = rng.random((400, dim)) * 1.6 - 0.3          # queries inside and outside the cloud
    Q[:5] = X[:5]                                   # exact hits -> zero distance rule
    knn = KNeighborsRegressor(n_neighbors=k, weights="distance").fit(X, y)
    dist_ref, idx_ref = knn.kneighbors(Q)
    pred_ref = knn.predict(Q)
    dist, idx = orc.knn_search(X, Q, k)
    assert np.array_equal(idx, idx_ref)
    assert np.array_equal(dist, dist_ref)
    assert np.array_equal(orc.idw_predict(dist, idx, y), pred_ref)


def test_knn_c_matches_numpy_twin():
    rng = np.random.default_rng(3)
    X = rng.random((3000, 3))
    Q = rng.random((50, 3))
    d0, i0 = orc.knn_search(X, Q, 26)
    d1, i1 = orc.knn_search_numpy(X, Q, 26)
    assert np.array_equal(i0, i1) and np.array_equal(d0, d1)


@pytest.mark.parametrize("k,D", [(8, 1), (8, 2), (26, 3)])
def test_interpolate_matches_torch_expression(k, D):
    # the reference's expression, export.py:467, evaluated by torch on the CPU
    rng = np.random.default_rng(k + D)
    N, Nc, T = 500, 200, 7
    data = rng.standard_normal((N, D, T)).astype(np.float32)
    idx = rng.integers(0, N, (Nc, k))
    w = rng.random((Nc, k))
    w /= w.sum(1, keepdims=True)
    wt, it, dt = pt.from_numpy(w), pt.from_numpy(idx), pt.from_numpy(data)
    ref = (wt[:, :, None, None] * dt[it]).sum(dim=1).numpy()
    out = orc.interpolate(w, idx, data, chunk_size=64)
    assert out.dtype == np.float64
    # torch's association order of the k-term sum depends on the tensor shape (probed: sequential for long rows,
    # 4 strided accumulators for short ones), so the fp64 results agree to rounding, not bit for bit
    np.testing.assert_allclose(out, ref, rtol=1e-13, atol=1e-14)


def test_export_weights_match_torch():
    # export.py:428-429
    rng = np.random.default_rng(0)
    d = rng.random((1000, 8))
    d[3, 0] = 0.0
    t = 1.0 / pt.clamp(pt.from_numpy(d), min=1e-12)
    t /= t.sum(axis=1, keepdim=True)
    assert np.array_equal(orc.export_weights(d), t.numpy())
    d26 = rng.random((1000, 26))
    t = 1.0 / pt.clamp(pt.from_numpy(d26), min=1e-12)
    t /= t.sum(axis=1, keepdim=True)
    np.testing.assert_allclose(orc.export_weights(d26), t.numpy(), rtol=1e-14)
