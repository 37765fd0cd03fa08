"""
Device k-nearest-neighbour index (host-side handle around ``s3_knn_*``).

Stands in for the two sklearn estimators the reference builds on the original CFD point cloud:
``KNeighborsRegressor(k, weights="distance")`` (sparseSpatialSampling/s_cube.py:161-163) and
``NearestNeighbors(k)`` (sparseSpatialSampling/export.py:120).
"""
import ctypes

import torch as pt

from . import _lib


def default_n_neighbors(n_dimensions: int) -> int:
    # reference: s_cube.py:161, export.py:117-118
    return 8 if n_dimensions == 2 else 26


class KnnIndex:
    def __init__(self, coordinates: pt.Tensor, values: pt.Tensor = None, device=None):
        """
        :param coordinates: ``[N, d]`` point cloud (any float dtype, CPU or CUDA); indexed as fp64
        :param values: optional ``[N]`` regression targets (the metric) for :meth:`predict`
        """
        _lib.require_cuda()
        self._lib = _lib.load()
        self.device = pt.device(device if device is not None else "cuda")
        coords = coordinates.detach().to(device=self.device, dtype=pt.float64).contiguous()
        assert coords.dim() == 2 and coords.size(1) in (2, 3), "coordinates must be [N, 2] or [N, 3]"
        vals = None
        if values is not None:
            vals = values.detach().to(device=self.device, dtype=pt.float64).contiguous()
            assert vals.dim() == 1 and vals.size(0) == coords.size(0)
        self.n, self.dim = coords.shape
        handle = ctypes.c_void_p()
        with pt.cuda.device(self.device):
            _lib.check(self._lib.s3_knn_build(_lib.ptr(coords), self.n, self.dim, _lib.ptr(vals), _lib.stream_ptr(),
                                              ctypes.byref(handle)))
        self._handle = handle

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            self._lib.s3_knn_free(h)
            self._handle = None

    @property
    def handle(self):
        return self._handle

    def _prep(self, query: pt.Tensor) -> pt.Tensor:
        q = query.detach().to(device=self.device, dtype=pt.float64).contiguous()
        if q.dim() == 1:
            q = q.unsqueeze(0)
        assert q.size(1) == self.dim
        return q

    def kneighbors(self, query: pt.Tensor, k: int):
        """Distances fp64 ``[Q, k]`` and indices int64 ``[Q, k]`` (ascending distance), on the device."""
        q = self._prep(query)
        idx = pt.empty((q.size(0), k), dtype=pt.int64, device=self.device)
        dist = pt.empty((q.size(0), k), dtype=pt.float64, device=self.device)
        with pt.cuda.device(self.device):
            _lib.check(self._lib.s3_knn_query(self._handle, _lib.ptr(q), q.size(0), k, _lib.ptr(idx), _lib.ptr(dist),
                                              _lib.stream_ptr()))
        return dist, idx

    def predict(self, query: pt.Tensor, k: int) -> pt.Tensor:
        """Inverse-distance weighted regression (sklearn ``weights='distance'`` semantics), fp64 ``[Q]``."""
        q = self._prep(query)
        pred = pt.empty((q.size(0),), dtype=pt.float64, device=self.device)
        with pt.cuda.device(self.device):
            _lib.check(self._lib.s3_knn_predict(self._handle, _lib.ptr(q), q.size(0), k, _lib.ptr(pred),
                                                _lib.stream_ptr()))
        return pred

    def tables(self, query: pt.Tensor, k: int, want_fp64: bool = True):
        """KNN cache of ``ExportData._build_knn_cache``: idx int32, w fp32 and (optionally) w fp64."""
        q = self._prep(query)
        idx = pt.empty((q.size(0), k), dtype=pt.int32, device=self.device)
        w32 = pt.empty((q.size(0), k), dtype=pt.float32, device=self.device)
        w64 = pt.empty((q.size(0), k), dtype=pt.float64, device=self.device) if want_fp64 else None
        with pt.cuda.device(self.device):
            _lib.check(self._lib.s3_knn_tables(self._handle, _lib.ptr(q), q.size(0), k, _lib.ptr(idx), _lib.ptr(w32),
                                               _lib.ptr(w64), _lib.stream_ptr()))
        return idx, w32, w64
