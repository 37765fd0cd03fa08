"""
Lowering of geometry objects to the parameter tables of the CUDA mask kernels, and the single-cell entry points
behind ``GeometryObject.check_cell``.

Table layout (``csrc/geometry.cuh``): header int32 ``[G, 4] = {type, keep_inside, parameter offset, n_extra}`` and
one flat fp64 parameter array.
"""
import torch as pt

import ctypes

import numpy as np

from .. import _lib
from .base import GEOM_CUSTOM

GEOM_STL = 7


class GeometryTable:
    def __init__(self, geometries: list, device):
        hdr, par = [], []
        self.custom = []          # indices of geometries without a device lowering (host check_cell)
        for i, g in enumerate(geometries):
            lowered = g.device_params() if hasattr(g, "device_params") else None
            if lowered is None:
                hdr.append([GEOM_CUSTOM, int(bool(g.keep_inside)), len(par), 0])
                self.custom.append(i)
                continue
            type_id, values, n_extra = lowered
            hdr.append([int(type_id), int(bool(g.keep_inside)), len(par), int(n_extra)])
            par.extend(values)
        if not par:
            par = [0.0]
        self.n = len(geometries)
        self.hdr = pt.tensor(hdr, dtype=pt.int32, device=device).contiguous()
        self.par = pt.tensor(par, dtype=pt.float64, device=device).contiguous()
        # closed triangulated surfaces take the tiled kernel (csrc/stl.cuh): bit mask of their positions and a HOST
        # copy of {parameter offset, n_triangles} per geometry for the launcher
        self.stl_geoms = 0
        for i, h in enumerate(hdr):
            if h[0] == GEOM_STL and i < 31:
                self.stl_geoms |= 1 << i
        self._stl_meta = np.ascontiguousarray([[h[2], h[3]] for h in hdr], dtype=np.int32)

    @property
    def stl_meta(self):
        """Host pointer of the int32 [n, 2] table {parameter offset, n_triangles} (NULL without STL geometries)."""
        return self._stl_meta.ctypes.data_as(ctypes.c_void_p) if self.stl_geoms else None


def nodes_invalid(geometries: list, nodes: pt.Tensor, refine_geometry: bool = False, only_geom: int = -1) -> pt.Tensor:
    """
    ``check_cell`` for ``n`` cells given as explicit nodes ``[n, n_nodes, d]`` -> bool ``[n]`` on the host.
    """
    _lib.require_cuda()
    lib = _lib.load()
    dev = pt.device("cuda", pt.cuda.current_device())
    table = GeometryTable(geometries, dev)
    assert not table.custom, "geometry without a device lowering: call its own check_cell"
    nd = nodes.detach().to(device=dev, dtype=pt.float64).contiguous()
    n, nn, dim = nd.shape
    out = pt.empty((n,), dtype=pt.uint8, device=dev)
    _lib.check(lib.s3_nodes_mask(_lib.ptr(nd), n, nn, dim, _lib.ptr(table.hdr), _lib.ptr(table.par), table.n,
                                 only_geom, int(bool(refine_geometry)), _lib.ptr(out), table.stl_geoms, table.stl_meta,
                                 _lib.stream_ptr()))
    return out.bool().cpu()


def nodes_inside(geometry, points: pt.Tensor) -> pt.Tensor:
    """Per-point inside mask of one geometry, bool ``[n]`` on the host."""
    _lib.require_cuda()
    lib = _lib.load()
    dev = pt.device("cuda", pt.cuda.current_device())
    table = GeometryTable([geometry], dev)
    p = points.detach().to(device=dev, dtype=pt.float64).contiguous()
    out = pt.empty((p.size(0),), dtype=pt.uint8, device=dev)
    _lib.check(lib.s3_points_inside(_lib.ptr(p), p.size(0), p.size(1), _lib.ptr(table.hdr), _lib.ptr(table.par), 0,
                                    _lib.ptr(out), table.stl_geoms, table.stl_meta, _lib.stream_ptr()))
    return out.bool().cpu()
