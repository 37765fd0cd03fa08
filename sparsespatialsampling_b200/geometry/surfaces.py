"""
Geometry plugins defined by a surface description: closed triangulated surfaces from STL files (3-D) and closed
polygons from coordinate lists (2-D).

Interfaces follow the reference classes ``GeometrySTL3D`` (sparseSpatialSampling/geometry/geometry_STL_3d.py:23-217)
and ``GeometryCoordinates2D`` (sparseSpatialSampling/geometry/coordinates_2d.py:16-157). The reference delegates the
point tests to VTK (``pyvista.select_enclosed_points``, ``check_surface=False``, tolerance 0.001) and shapely
(``Point.within`` = strict interior); neither library is available offline, so the tests implemented in
``csrc/geometry.cuh`` are documented restatements:

* STL: a point is inside if it lies within ``0.001 * bounding-box diagonal`` of the surface (VTK's tolerance
  convention; this is what makes the reference's on-surface test cells count as inside), otherwise by the parity of
  the crossings of a +x ray;
* polygon: even-odd rule, points on the boundary are outside.

Parity with the reference is pinned only by the reference's own unit tests for these two classes (restated in
tests/test_geometry_surfaces_gpu.py); ``reduce_by`` (VTK decimation) and the pymeshfix auto-repair are not available:
the STL file must already be closed and manifold.
"""
import logging
import struct
from typing import Union

import numpy as np
import torch as pt

from .base import GeometryObject, GEOM_STL, GEOM_POLY2D
from .analytic import CubeGeometry

logger = logging.getLogger(__name__)


def read_stl(path: str) -> np.ndarray:
    """Triangles ``[n, 3, 3]`` (fp64) of a binary or ASCII STL file."""
    with open(path, "rb") as f:
        raw = f.read()
    if len(raw) >= 84:
        n = struct.unpack("<I", raw[80:84])[0]
        if 84 + 50 * n == len(raw):
            rec = np.frombuffer(raw, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]), count=n,
                                offset=84)
            return rec["v"].astype(np.float64)
    tris, cur = [], []
    for line in raw.decode("ascii", errors="ignore").splitlines():
        parts = line.split()
        if len(parts) == 4 and parts[0] == "vertex":
            cur.append([float(parts[1]), float(parts[2]), float(parts[3])])
            if len(cur) == 3:
                tris.append(cur)
                cur = []
    if not tris:
        raise ValueError(f"Could not read any triangle from STL file {path}.")
    return np.asarray(tris, dtype=np.float64)


def is_closed_manifold(triangles: np.ndarray) -> bool:
    """Every undirected edge is shared by exactly two triangles (vertices welded by exact coordinates)."""
    verts, inv = np.unique(triangles.reshape(-1, 3), axis=0, return_inverse=True)
    f = inv.reshape(-1, 3)
    edges = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]], axis=0)
    edges.sort(axis=1)
    _, counts = np.unique(edges, axis=0, return_counts=True)
    return bool((counts == 2).all())


class _BoundedGeometry(GeometryObject):
    """Shared pieces of the two classes: bounding box, main width, centre, bounding-box pre-check."""

    def _finish_bounds(self, lower, upper):
        self._lower_bound = [float(v) for v in lower]
        self._upper_bound = [float(v) for v in upper]
        self._main_width = max(abs(u - l) for l, u in zip(self._lower_bound, self._upper_bound))
        self._center = (pt.tensor(self._lower_bound, dtype=pt.float64) +
                        pt.tensor(self._upper_bound, dtype=pt.float64)) / 2.0
        self._bbox = CubeGeometry(f"{self._name}_bbox", self._keep_inside, self._lower_bound, self._upper_bound)

    def pre_check_cell(self, cell_nodes: pt.Tensor, refine_geometry: bool = False) -> bool:
        """``check_cell`` against the bounding box only (geometry_STL_3d.py:105-124, coordinates_2d.py:75-94)."""
        return self._bbox.check_cell(cell_nodes, refine_geometry)

    @property
    def type(self) -> str:
        return self._type

    @property
    def main_width(self) -> float:
        return self._main_width

    @property
    def center(self) -> pt.Tensor:
        return self._center


class GeometrySTL3D(_BoundedGeometry):
    __short_description__ = "usage of STL files for geometries (3D)"

    def __init__(self, name: str, keep_inside: bool, path_stl_file: str, refine: bool = False,
                 min_refinement_level: int = None, reduce_by: Union[int, float] = 0):
        if reduce_by < 0:
            logger.warning(f"Found invalid negative value for 'reduce_by' of {reduce_by}. Disabling compression.")
            reduce_by = 0
        if reduce_by > 0:
            raise NotImplementedError("reduce_by > 0 needs VTK's quadric decimation (pyvista), which is not available; "
                                      "pass an already decimated STL file.")
        super().__init__(name, keep_inside, refine, min_refinement_level)
        self._type = "STL"
        self._pwd = path_stl_file
        self._triangles = read_stl(path_stl_file)
        pts = self._triangles.reshape(-1, 3)
        self._finish_bounds(pts.min(0), pts.max(0))
        diag = float(np.linalg.norm(pts.max(0) - pts.min(0)))
        self._tolerance = 0.001 * diag           # pyvista select_enclosed_points(tolerance=0.001) * bbox diagonal
        self._check_geometry()

    def _check_geometry(self) -> None:
        if not is_closed_manifold(self._triangles):
            raise RuntimeError(f"Expected an STL file with a closed and manifold surface for geometry {self.name}; "
                               f"automatic repair (pymeshfix) is not available.")
        n_points = np.unique(self._triangles.reshape(-1, 3), axis=0).shape[0]
        if n_points > 5e4:
            logger.warning(f"STL file contains {n_points} points; this slows down the geometry masks. "
                           f"Consider a coarser STL file (< 5e4 points).")

    TILE = 128                                    # triangles per shared-memory tile of the device kernel (csrc/stl.cuh)

    def device_params(self):
        """lo[3], hi[3], tolerance, the triangles in Morton order of their centroids (n_tri * 9) and one bounding box per
        tile of 128 consecutive triangles (n_tiles * 6): spatially compact tiles are what lets a CTA skip most of them."""
        tri = self._tiled_triangles()
        n_tiles = (tri.shape[0] + self.TILE - 1) // self.TILE
        boxes = np.empty((n_tiles, 6), dtype=np.float64)
        for t in range(n_tiles):
            v = tri[t * self.TILE:(t + 1) * self.TILE].reshape(-1, 3)
            boxes[t, :3], boxes[t, 3:] = v.min(0), v.max(0)
        par = (self._lower_bound + self._upper_bound + [self._tolerance] + tri.reshape(-1).tolist() +
               boxes.reshape(-1).tolist())
        return GEOM_STL, par, int(tri.shape[0])

    def _tiled_triangles(self) -> np.ndarray:
        tri = np.asarray(self._triangles, dtype=np.float64).reshape(-1, 3, 3)
        lo, hi = np.asarray(self._lower_bound), np.asarray(self._upper_bound)
        cen = (tri.mean(axis=1) - lo) / np.maximum(hi - lo, 1e-300)
        q = np.clip((cen * 1023.0).astype(np.int64), 0, 1023)             # 10 bits per axis
        code = np.zeros(tri.shape[0], dtype=np.int64)
        for b in range(10):
            for a in range(3):
                code |= ((q[:, a] >> b) & 1) << (3 * b + a)
        return tri[np.argsort(code, kind="stable")]


class GeometryCoordinates2D(_BoundedGeometry):
    __short_description__ = "2D coordinates for geometries"

    def __init__(self, name: str, keep_inside: bool, coordinates: Union[list, np.ndarray], refine: bool = False,
                 min_refinement_level: int = None):
        super().__init__(name, keep_inside, refine, min_refinement_level)
        self._type = "coord_2D"
        v = np.asarray(coordinates, dtype=np.float64).reshape(-1, 2)
        if v.shape[0] > 1 and np.array_equal(v[0], v[-1]):
            v = v[:-1]                            # a ring repeats its first point; the polygon closes implicitly
        self._vertices = v
        self._check_geometry()
        self._finish_bounds(v.min(0), v.max(0))

    def _check_geometry(self) -> None:
        assert self._vertices.shape[0] >= 3, (f"Expected at least three coordinates forming an enclosed area for "
                                              f"geometry {self.name}.")
        x, y = self._vertices[:, 0], self._vertices[:, 1]
        area = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
        assert area > 0, f"The coordinates of geometry {self.name} do not enclose an area."

    def device_params(self):
        par = self._lower_bound + self._upper_bound + self._vertices.reshape(-1).tolist()
        return GEOM_POLY2D, par, int(self._vertices.shape[0])
