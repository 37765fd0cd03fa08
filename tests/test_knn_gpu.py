"""GPU parity: device KNN (s3_knn_*) against the CPU oracle -- indices bit-exact, values bit-exact."""
import numpy as np
import pytest
import torch as pt

from oracle import s3_oracle as orc

pytestmark = pytest.mark.gpu


def _cloud(n, dim, seed, clustered=False):
    rng = np.random.default_rng(seed)
    X = rng.random((n, dim))
    if clustered:  # anisotropic, CFD-like clustering towards a wall
        X[:, 1] = X[:, 1] ** 3 * 0.41
        X[:, 0] *= 2.2
    return X


@pytest.mark.parametrize("n,dim,k,clustered", [
    (20000, 2, 8, True), (30000, 3, 26, False), (1000, 2, 8, False), (33, 3, 9, False), (50000, 2, 1, False),
    (40, 2, 8, False), (100000, 3, 26, True),
])
def test_kneighbors_bit_exact(cuda, n, dim, k, clustered):
    from sparsespatialsampling_b200.knn import KnnIndex
    X = _cloud(n, dim, n + dim, clustered)
    rng = np.random.default_rng(7)
    lo, hi = X.min(0), X.max(0)
    Q = lo + (rng.random((3000, dim)) * 1.8 - 0.4) * (hi - lo)   # inside, near and far outside the cloud
    Q[:10] = X[:10]                                               # exact hits
    index = KnnIndex(pt.from_numpy(X))
    dist, idx = index.kneighbors(pt.from_numpy(Q), k)
    d_ref, i_ref = orc.knn_search(X, Q, k)
    assert np.array_equal(idx.cpu().numpy(), i_ref)
    assert np.array_equal(dist.cpu().numpy(), d_ref)


@pytest.mark.parametrize("dim,k", [(2, 8), (3, 26), (3, 5)])
def test_idw_predict_bit_exact(cuda, dim, k):
    # KNeighborsRegressor(weights="distance").predict as called at s_cube.py:224
    from sparsespatialsampling_b200.knn import KnnIndex
    X = _cloud(40000, dim, 11 * dim + k)
    rng = np.random.default_rng(5)
    y = rng.random(X.shape[0]) * 3.0
    Q = rng.random((5000, dim)) * 1.4 - 0.2
    Q[:20] = X[100:120]                                           # zero distance -> indicator weights
    index = KnnIndex(pt.from_numpy(X), pt.from_numpy(y))
    pred = index.predict(pt.from_numpy(Q), k).cpu().numpy()
    ref = orc.knn_predict(X, y, Q, k)
    assert np.array_equal(pred, ref)


def test_idw_predict_matches_sklearn_directly(cuda):
    from sklearn.neighbors import KNeighborsRegressor
    from sparsespatialsampling_b200.knn import KnnIndex
    X = _cloud(25000, 3, 99)
    y = np.random.default_rng(1).random(25000)
    Q = np.random.default_rng(2).random((2000, 3))
    ref = KNeighborsRegressor(n_neighbors=26, weights="distance").fit(X, y).predict(Q)
    pred = KnnIndex(pt.from_numpy(X), pt.from_numpy(y)).predict(pt.from_numpy(Q), 26).cpu().numpy()
    assert np.array_equal(pred, ref)


@pytest.mark.parametrize("dim,k", [(2, 8), (3, 26)])
def test_export_tables(cuda, dim, k):
    # ExportData._build_knn_cache, export.py:403-444
    from sparsespatialsampling_b200.knn import KnnIndex
    X = _cloud(30000, dim, 17)
    Q = np.random.default_rng(3).random((4000, dim))
    Q[0] = X[5]
    idx, w32, w64 = KnnIndex(pt.from_numpy(X)).tables(pt.from_numpy(Q), k)
    d_ref, i_ref = orc.knn_search(X, Q, k)
    w_ref = orc.export_weights(d_ref)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), i_ref)
    if k == 8:
        assert np.array_equal(w64.cpu().numpy(), w_ref)
    else:
        np.testing.assert_allclose(w64.cpu().numpy(), w_ref, rtol=1e-14)
    np.testing.assert_allclose(w32.cpu().numpy(), w_ref.astype(np.float32), rtol=2e-7)


def test_duplicate_points_ties_by_index(cuda):
    from sparsespatialsampling_b200.knn import KnnIndex
    rng = np.random.default_rng(0)
    X = rng.random((5000, 2))
    X[2500:] = X[:2500]                         # every point twice -> exact distance ties
    Q = rng.random((500, 2))
    dist, idx = KnnIndex(pt.from_numpy(X)).kneighbors(pt.from_numpy(Q), 8)
    d_ref, i_ref = orc.knn_search(X, Q, 8)
    assert np.array_equal(idx.cpu().numpy(), i_ref)
    assert np.array_equal(dist.cpu().numpy(), d_ref)
