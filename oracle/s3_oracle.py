"""
TEST INFRASTRUCTURE ONLY -- CPU oracle of the S^3 hot path. Never imported by the product package.

A plain numpy (+ one small C file, oracle/knn_oracle.c) restatement of the reference's algorithm for the path
named in BASELINE.json: KNN / inverse-distance prediction, gain, geometry masks, the refinement loop and the
export-stage interpolation. Every function cites the reference file:line (relative to the reference repository
JanisGeise/sparseSpatialSampling v1.0.0) or the third-party routine it follows.

Pinning (see DESIGN.md "Oracle"): the functions here are checked
  * against scikit-learn 1.9.0 itself (present in this image) for KNN indices / distances / predictions
    (tests/test_oracle.py), and
  * against outputs of the reference run in the build container, committed as fixtures under tests/golden/
    by tests/golden/make_golden.py (which imports /root/reference; it is not needed at test time).
"""
import ctypes
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libs3oracle.so")
_clib = None


def _load_c():
    global _clib
    if _clib is None:
        if not os.path.exists(_SO):
            try:
                from . import build as _b
            except ImportError:  # imported as a top-level module
                import importlib.util
                spec = importlib.util.spec_from_file_location("s3_oracle_build", os.path.join(_HERE, "build.py"))
                _b = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(_b)
            _b.build()
        lib = ctypes.CDLL(_SO)
        lib.s3o_knn.restype = None
        lib.s3o_knn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        _clib = lib
    return _clib


# --------------------------------------------------------------------------------------------- KNN
def knn_search(points: np.ndarray, query: np.ndarray, k: int):
    """
    Exact k nearest neighbours, ascending distance (ties: smaller index).
    Restates sklearn.neighbors.KDTree.query as used via kneighbors() (s_cube.py:224; export.py:425,438).
    Returns (dist fp64 [Q,k], idx int64 [Q,k]).
    """
    X = np.ascontiguousarray(points, dtype=np.float64)
    Q = np.ascontiguousarray(query, dtype=np.float64)
    if Q.ndim == 1:
        Q = Q[None, :]
    nq = Q.shape[0]
    idx = np.empty((nq, k), dtype=np.int64)
    dist = np.empty((nq, k), dtype=np.float64)
    lib = _load_c()
    lib.s3o_knn(X.ctypes.data, X.shape[0], X.shape[1], Q.ctypes.data, nq, k, idx.ctypes.data, dist.ctypes.data)
    return dist, idx


def knn_search_numpy(points: np.ndarray, query: np.ndarray, k: int):
    """Pure-numpy twin of :func:`knn_search` (small inputs only); used to cross-check the C file."""
    X = np.asarray(points, dtype=np.float64)
    Q = np.atleast_2d(np.asarray(query, dtype=np.float64))
    dist = np.empty((Q.shape[0], k))
    idx = np.empty((Q.shape[0], k), dtype=np.int64)
    ar = np.arange(X.shape[0])
    for i in range(Q.shape[0]):
        rd = np.zeros(X.shape[0])
        for j in range(X.shape[1]):
            t = Q[i, j] - X[:, j]
            rd = rd + t * t
        order = np.lexsort((ar, rd))[:k]
        idx[i] = order
        dist[i] = np.sqrt(rd[order])
    return dist, idx


def _pairwise8_rows(a: np.ndarray) -> np.ndarray:
    """
    Row sums in numpy's pairwise-summation order for a contiguous inner axis of length n <= 128
    (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum): what np.sum(a, axis=1) evaluates.
    """
    n = a.shape[1]
    if n < 8:
        res = np.zeros(a.shape[0])
        for i in range(n):
            res = res + a[:, i]
        return res
    r = [a[:, j].copy() for j in range(8)]
    n8 = n - (n % 8)
    for i in range(8, n8, 8):
        for j in range(8):
            r[j] = r[j] + a[:, i + j]
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    for i in range(n8, n):
        res = res + a[:, i]
    return res


def idw_predict(dist: np.ndarray, idx: np.ndarray, values: np.ndarray) -> np.ndarray:
    """
    KNeighborsRegressor(weights="distance").predict (sklearn/neighbors/_regression.py predict +
    sklearn/neighbors/_base.py _get_weights), the call at s_cube.py:224,328,372.
    """
    y = np.asarray(values, dtype=np.float64)
    with np.errstate(divide="ignore"):
        w = 1.0 / dist
    inf_mask = np.isinf(w)
    inf_row = inf_mask.any(axis=1)
    w[inf_row] = inf_mask[inf_row]
    num = _pairwise8_rows(y[idx] * w)
    den = _pairwise8_rows(w)
    return num / den


def knn_predict(points, values, query, k):
    d, i = knn_search(points, query, k)
    return idw_predict(d, i, values)


def export_weights(dist: np.ndarray) -> np.ndarray:
    """
    ExportData._build_knn_cache (export.py:428-429): w = 1/clamp(dist, 1e-12); w /= w.sum(1).
    The row sum uses torch's fp64 sum(dim=1) order for 8 columns, (((a0+a4)+(a1+a5))+(a2+a6))+(a3+a7)
    (probed on the build host); other widths are summed pairwise and compared with a tolerance.
    """
    w = 1.0 / np.maximum(dist, 1e-12)
    if w.shape[1] == 8:
        s = (((w[:, 0] + w[:, 4]) + (w[:, 1] + w[:, 5])) + (w[:, 2] + w[:, 6])) + (w[:, 3] + w[:, 7])
    else:
        s = _pairwise8_rows(w)
    return w / s[:, None]


# ------------------------------------------------------------------------------------ interpolation
def interpolate(weights: np.ndarray, idx: np.ndarray, data: np.ndarray, chunk_size: int = 100000) -> np.ndarray:
    """
    interpolate_data (export.py:446-468): out[c] = sum_j w[c, j] * data[idx[c, j]], fp64 result,
    products first, then a sequential sum over the neighbour axis.
    """
    nc, k = idx.shape
    out = np.empty((nc,) + data.shape[1:], dtype=np.float64)
    for s in range(0, nc, chunk_size):
        e = min(s + chunk_size, nc)
        g = data[idx[s:e]]                                   # [chunk, k, ...]
        wv = weights[s:e].reshape((e - s, k) + (1,) * (data.ndim - 1))
        prod = wv * g
        acc = prod[:, 0].astype(np.float64)
        for j in range(1, k):
            acc = acc + prod[:, j]
        out[s:e] = acc
    return out
