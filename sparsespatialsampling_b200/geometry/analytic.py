"""
Analytic geometry plugins: box, sphere/circle, cylinder/cone, triangle, prism, tetrahedron, pyramid.

Constructor signatures, argument checks, ``main_width`` / ``center`` and the inside/outside semantics follow the
reference classes (sparseSpatialSampling/geometry/{cube,sphere,cylinder,triangle,prism,tetrahedron,pyramid}_geometry.py).
The point-in-shape arithmetic itself lives in ``csrc/geometry.cuh``; each class only packs its parameter block.
All set-up arithmetic that feeds the kernel (axis vectors, norms, face normals) is done here with the same torch
CPU operators the reference uses, so the parameters are bit-identical to the reference's attributes.
"""
from typing import List, Union

import torch as pt

from .base import (GeometryObject, GEOM_CUBE, GEOM_SPHERE, GEOM_CYLINDER, GEOM_TRIANGLE, GEOM_PRISM, GEOM_TETRA,
                   GEOM_PYRAMID)

F64 = pt.float64


def _f64_list(t) -> list:
    return [float(v) for v in pt.as_tensor(t, dtype=F64).flatten().tolist()]


class CubeGeometry(GeometryObject):
    __short_description__ = "rectangles (2D) or cubes (3D)"

    def __init__(self, name: str, keep_inside: bool, lower_bound: list, upper_bound: list, refine: bool = False,
                 min_refinement_level: int = None):
        super().__init__(name, keep_inside, refine, min_refinement_level)
        self._lower_bound = lower_bound
        self._upper_bound = upper_bound
        self._type = "cube"
        self._check_geometry()
        # cube_geometry.py:136-150
        self._main_width = max(abs(u - l) for l, u in zip(self._lower_bound, self._upper_bound))
        self._center = (pt.tensor(self._lower_bound, dtype=F64) + pt.tensor(self._upper_bound, dtype=F64)) / 2.0

    def _check_geometry(self) -> None:
        assert self._lower_bound, "Found empty list for the lower bound. Please provide values for the lower bound."
        assert self._upper_bound, "Found empty list for the upper bound. Please provide values for the upper bound."
        assert len(self._lower_bound) == len(self._upper_bound), (
            f"Found {len(self._lower_bound)} values for the lower bound but {len(self._upper_bound)} values for the "
            f"upper bound for geometry {self.name}.")
        for i, (lo, up) in enumerate(zip(self._lower_bound, self._upper_bound)):
            assert lo < up, (f"Lower bound {lo} at position {i} is not smaller than the upper bound {up} "
                             f"for geometry {self.name}.")

    def _check_dimensions(self, cell_nodes):
        assert cell_nodes.size(-1) == len(self._lower_bound), (
            f"Number of dimensions of the cell does not match the number of given bounds. Expected "
            f"{cell_nodes.size(-1)} values, found {len(self._lower_bound)} for geometry {self.name}.")

    def device_params(self):
        return GEOM_CUBE, [float(v) for v in self._lower_bound] + [float(v) for v in self._upper_bound], 0

    @property
    def type(self) -> str:
        return self._type

    @property
    def main_width(self) -> float:
        return self._main_width

    @property
    def center(self) -> pt.Tensor:
        return self._center


class SphereGeometry(GeometryObject):
    __short_description__ = "circles (2D) or spheres (3D)"

    def __init__(self, name: str, keep_inside: bool, position: list, radius: Union[int, float], refine: bool = False,
                 min_refinement_level: int = None):
        super().__init__(name, keep_inside, refine, min_refinement_level)
        self._position = position
        self._radius = radius
        self._type = "sphere"
        # sphere_geometry.py:121-138 (sic: the main width is the radius)
        self._main_width = float(self._radius)
        self._center = pt.tensor(self._position, dtype=F64)
        self._check_geometry()

    def _check_geometry(self) -> None:
        assert self._position, "Found empty list for the position. Please provide values for the position."
        assert isinstance(self._radius, (int, float)), (f"Expected the type of radius to be Union[int, float], got "
                                                        f"{type(self._radius)} for geometry {self.name} instead.")
        assert self._radius > 0, f"Expected a radius larger than zero but found a value of {self._radius}."

    def _check_dimensions(self, cell_nodes):
        assert cell_nodes.size(1) == len(self._position), (
            f"Number of dimensions of the cell does not match the number of dimensions for the position. Expected "
            f"{cell_nodes.size(-1)} values, found {len(self._position)} for geometry {self.name}.")

    def device_params(self):
        return GEOM_SPHERE, [float(v) for v in self._position] + [float(self._radius)], 0

    @property
    def type(self) -> str:
        return self._type

    @property
    def main_width(self) -> float:
        return self._main_width

    @property
    def center(self) -> pt.Tensor:
        return self._center


class CylinderGeometry3D(GeometryObject):
    __short_description__ = "cylinders, conical objects and cones (3D)"

    def __init__(self, name: str, keep_inside: bool, position: List[Union[list, tuple]],
                 radius: Union[int, float, list, tuple], refine: bool = False, min_refinement_level: int = None):
        super().__init__(name, keep_inside, refine, min_refinement_level)
        self._position = position
        self._radius = radius
        self._type = "cylinder"
        self._check_geometry()
        # cylinder_geometry.py:52-58: the end points are rounded to float32, the axis is their fp32 difference
        self._position = pt.tensor(self._position).float()
        self._axis = (self._position[1, :] - self._position[0, :]).type(F64)
        self._norm = self._axis.norm()
        r_max = max(self._radius) if isinstance(self._radius, (list, tuple)) else self._radius
        self._main_width = max(r_max, self._axis.norm().item())
        self._center = self._position.mean(0)

    def _check_geometry(self) -> None:
        assert self._position, "Found empty list for the position. Please provide values for the positions."
        assert len(self._position) == 2, (f"Expected exactly two entries for the position but found "
                                          f"{len(self._position)} entries.")
        assert self._position[0] != self._position[1], "Expected two different positions, a cylinder of length zero " \
                                                       "is invalid."
        assert isinstance(self._radius, (int, float, list, tuple)), (
            f"Expected the type of radius to be Union[int, float, list, tuple], got {type(self._radius)} for "
            f"geometry {self.name} instead.")
        if isinstance(self._radius, (int, float)):
            assert self._radius > 0, f"Expected a radius larger than zero but found a value of {self._radius}."
        else:
            assert len(self._radius) == 2, f"Expected two values for the radii but found {len(self._radius)}."
            assert self._radius[0] >= 0 and self._radius[1] >= 0, (f"Expected all radii >= 0 but found a values of "
                                                                   f"{self._radius}.")
            assert (self._radius[0] == self._radius[1]) == 0, (f"Both values for the radii can't be zero. At least "
                                                               f"one radius has to be > 0 but found values of "
                                                               f"{self._radius}.")

    def device_params(self):
        cone = not isinstance(self._radius, (int, float))
        r0, r1 = (self._radius[0], self._radius[1]) if cone else (self._radius, self._radius)
        par = _f64_list(self._position[0, :]) + _f64_list(self._axis) + [float(self._norm.item()), float(r0),
                                                                         float(r1), 1.0 if cone else 0.0]
        return GEOM_CYLINDER, par, 0

    @property
    def type(self) -> str:
        return self._type

    @property
    def main_width(self) -> float:
        return self._main_width

    @property
    def center(self) -> pt.Tensor:
        return self._center


class TriangleGeometry(GeometryObject):
    __short_description__ = "triangles (2D)"

    def __init__(self, name: str, keep_inside: bool, points: Union[list, pt.Tensor], refine: bool = False,
                 min_refinement_level: int = None):
        super().__init__(name, keep_inside, refine, min_refinement_level)
        self._type = "triangle"
        assert isinstance(points, (list, tuple, pt.Tensor)), (f"Expected the points to be a list or pt.Tensor, but "
                                                              f"found type {type(points)} instead.")
        # python lists are read as fp64 (the reference runs with torch's default dtype set to float64)
        self._points = [p.type(F64) if isinstance(p, pt.Tensor) else pt.tensor(p, dtype=F64) for p in points]
        self._check_geometry()
        stacked = pt.stack(self._points, dim=0)
        # triangle_geometry.py:180-199
        self._main_width = (stacked.max(0).values - stacked.min(0).values).abs().max().item()
        self._center = stacked.mean(0)

    def _check_geometry(self) -> None:
        assert len(self._points) == 3, f"Expected 3 points, but found {len(self._points)} points instead."
        assert all(len(p) == 2 for p in self._points), ("All given coordinates have to contain exactly 2 entries "
                                                        "with the x- and y-coordinates.")
        a = self._points[1] - self._points[0]
        b = self._points[2] - self._points[0]
        area = 0.5 * abs(a[0] * b[1] - a[1] * b[0])
        assert area > 0, f"The area of the triangle has to be larger than zero. Found an area of {area}."

    def device_params(self):
        return GEOM_TRIANGLE, _f64_list(pt.stack(self._points, 0)), 0

    def check_triangle(self, vertices: pt.Tensor) -> pt.Tensor:
        """Per-node inside mask (used by the reference's prism through ``check_triangle``)."""
        from .device import nodes_inside
        return nodes_inside(self, vertices)

    @property
    def type(self) -> str:
        return self._type

    @property
    def main_width(self) -> float:
        return self._main_width

    @property
    def center(self) -> pt.Tensor:
        return self._center


class PrismGeometry3D(GeometryObject):
    __short_description__ = "prisms (3D)"

    def __init__(self, name: str, keep_inside: bool, positions: List[List[Union[list, tuple]]], refine: bool = False,
                 min_refinement_level: int = None):
        super().__init__(name, keep_inside, refine, min_refinement_level)
        self._type = "prism"
        self._positions = positions
        self._check_geometry()
        self._positions = [pt.tensor(tri, dtype=F64) for tri in self._positions]
        # prism_geometry.py:44-66: extrusion axis = first point of the 2nd minus first point of the 1st triangle
        self._axis = (self._positions[1][0] - self._positions[0][0]).type(F64)
        self._norm = self._axis.norm()
        self._dim = pt.where(self._axis == 0)[0]
        assert len(self._dim) == 2, "The specified triangles are not aligned along a coordinate direction."
        assert pt.allclose(self._positions[0][:, self._dim], self._positions[1][:, self._dim]), \
            "The specified triangles are not aligned along a coordinate direction."
        self._triangles = [
            TriangleGeometry(f"{name}_first", keep_inside=True, points=self._positions[0][:, self._dim]),
            TriangleGeometry(f"{name}_second", keep_inside=True, points=self._positions[1][:, self._dim])]
        self._main_width = max(self._axis.norm().item(), max(t.main_width for t in self._triangles))
        self._center = self._compute_center()

    def _check_geometry(self) -> None:
        assert self._positions, "Found empty list for the positions. Please provide values for the prism."
        assert len(self._positions) == 2, (f"Expected exactly two triangles for the prism but found "
                                           f"{len(self._positions)} entries.")
        assert all(len(tri) == 3 for tri in self._positions), "Each triangle must have exactly 3 vertices."

    def _compute_center(self) -> pt.Tensor:
        # prism_geometry.py:176-197
        tri_centers = pt.cat([t.center.unsqueeze(-1) for t in self._triangles], -1).mean(1)
        ax_dim = self._axis.nonzero()[0]
        if len(ax_dim) > 1:
            raise NotImplementedError("The triangles are not aligned along a coordinate axis, which is currently not"
                                      " supported.")
        ax = ax_dim.item()
        avg = (self._positions[1][0, ax] + self._positions[0][0, ax]) / 2
        out = pt.zeros((3,), dtype=self._axis.dtype)
        out[ax] = avg
        out[self._dim] = tri_centers
        return out

    def device_params(self):
        par = (_f64_list(self._positions[0][0]) + _f64_list(self._axis) + [float(self._norm.item())] +
               [float(self._dim[0].item()), float(self._dim[1].item())] +
               _f64_list(self._positions[0][:, self._dim]))
        return GEOM_PRISM, par, 0

    @property
    def type(self) -> str:
        return self._type

    @property
    def main_width(self) -> float:
        return self._main_width

    @property
    def center(self) -> pt.Tensor:
        return self._center


class TetrahedronGeometry3D(GeometryObject):
    __short_description__ = "tetrahedrons (3D)"

    def __init__(self, name: str, keep_inside: bool, positions: Union[List[Union[list, tuple]], pt.Tensor],
                 refine: bool = False, min_refinement_level: int = None):
        super().__init__(name, keep_inside, refine, min_refinement_level)
        self._type = "tetrahedron"
        self._positions = positions
        self._normals = None
        self._check_geometry()
        if not isinstance(self._positions, pt.Tensor):
            self._positions = pt.tensor(self._positions, dtype=F64)
        else:
            self._positions = self._positions.type(F64)
        # volume = |det([P, 1])| / 6 must not vanish (tetrahedron_geometry.py:56-61)
        vol = pt.det(pt.cat([self._positions, pt.ones((4, 1), dtype=F64)], dim=1)).abs() / 6
        assert vol > 0, "The tetrahedron provided has a volume of zero."
        self._compute_normals()
        self._main_width = (self._positions.max(dim=0).values - self._positions.min(dim=0).values).max().item()
        self._center = self._positions.mean(dim=0)

    def _compute_normals(self) -> None:
        # tetrahedron_geometry.py:70-104: one normal per point index p, flipped to point towards the centroid.
        # (n4 uses D - C as its second edge, exactly as the reference does.)
        P = self._positions
        centroid = P.mean(dim=0)
        n1 = pt.cross(P[1] - P[0], P[2] - P[0], dim=0).unsqueeze(-1)
        n2 = pt.cross(P[1] - P[0], P[3] - P[0], dim=0).unsqueeze(-1)
        n3 = pt.cross(P[2] - P[0], P[3] - P[0], dim=0).unsqueeze(-1)
        n4 = pt.cross(P[2] - P[1], P[3] - P[2], dim=0).unsqueeze(-1)
        normals = pt.cat([n1, n2, n3, n4], dim=1)
        check = [pt.dot(centroid - P[p, :], normals[:, p]) for p in range(4)]
        normals[:, pt.where(pt.tensor(check) < 0)[0]] *= -1
        self._normals = normals

    def _check_geometry(self) -> None:
        if isinstance(self._positions, list):
            assert self._positions, "Found empty list for the positions. Please provide values for the tetrahedron."
        else:
            assert isinstance(self._positions, pt.Tensor), (f"Expected positions to be a list or tensor but found "
                                                            f"type {type(self._positions)}.")
        assert len(self._positions) == 4, (f"Expected exactly four points for the tetrahedron but found "
                                           f"{len(self._positions)} entries.")
        assert all(len(p) == 3 for p in self._positions), "Each point of the tetrahedron needs three coordinates."

    def device_params(self):
        # positions [4][3] then normal of point p as a contiguous triple
        return GEOM_TETRA, _f64_list(self._positions) + _f64_list(self._normals.t().contiguous()), 0

    def check_tetrahedron(self, vertices: pt.Tensor) -> pt.Tensor:
        from .device import nodes_inside
        return nodes_inside(self, vertices)

    @property
    def type(self) -> str:
        return self._type

    @property
    def main_width(self) -> float:
        return self._main_width

    @property
    def center(self) -> pt.Tensor:
        return self._center


class PyramidGeometry3D(GeometryObject):
    __short_description__ = "square pyramids (3D)"

    def __init__(self, name: str, keep_inside: bool, nodes: List[Union[list, tuple]], refine: bool = False,
                 min_refinement_level: int = None):
        super().__init__(name, keep_inside, refine, min_refinement_level)
        self._type = "pyramid"
        self._nodes = nodes
        self._check_geometry()
        self._nodes = pt.tensor(self._nodes, dtype=F64)
        self._split_into_tetrahedra()
        self._main_width = max(t.main_width for t in self._tets)
        self._center = pt.cat([t.center.unsqueeze(-1) for t in self._tets], -1).mean(1)

    def _check_geometry(self) -> None:
        assert len(self._nodes) == 5, f"Expected exactly five vertices for the pyramid but found {len(self._nodes)} " \
                                      f"vertices."
        for i, v in enumerate(self._nodes):
            assert isinstance(v, (list, tuple)), f"Expected each vertex to be a list or tuple but found type " \
                                                 f"{type(v)} for vertex no. {i}."
            assert len(v) == 3, f"Expected three coordinates per vertex but found {len(v)} for entry {i}."

    def _split_into_tetrahedra(self) -> None:
        # pyramid_geometry.py:52-154: the base is the plane through three vertices that contains the most vertices,
        # the apex the vertex farthest from it; the base is cut along its longest diagonal.
        N = self._nodes
        best, base_n, base_p = 0, None, None
        for i in range(5):
            for j in range(i + 1, 5):
                for k in range(j + 1, 5):
                    n = pt.cross(N[j] - N[i], N[k] - N[i], dim=0)
                    if n.norm() < 1e-12:
                        continue
                    n = n / n.norm()
                    inliers = (abs((N - N[i]) @ n) < 1e-6).sum()
                    if inliers > best:
                        best, base_n, base_p = inliers, n, N[i]
        if base_n is None:
            raise RuntimeError("No valid plane detected: the vertices may be collinear.")
        self._apex_idx = pt.argmax(abs((N - base_p) @ base_n)).item()
        base_idx = [i for i in range(5) if i != self._apex_idx]
        pts = N[base_idx]
        d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
        d2.fill_diagonal_(-float("inf"))
        i, j = pt.nonzero(d2 == d2.max(), as_tuple=True)
        self._diagonal_idx = (base_idx[i[0].item()], base_idx[j[0].item()])
        self._off_diagonal = [i for i in base_idx if i not in self._diagonal_idx]
        idx1 = [self._diagonal_idx[0], self._off_diagonal[0], self._diagonal_idx[1], self._apex_idx]
        idx2 = [self._diagonal_idx[1], self._off_diagonal[1], self._diagonal_idx[0], self._apex_idx]
        self._tets = [TetrahedronGeometry3D("tet0", self._keep_inside, N[idx1]),
                      TetrahedronGeometry3D("tet1", self._keep_inside, N[idx2])]

    def device_params(self):
        par = []
        for t in self._tets:
            par += t.device_params()[1]
        return GEOM_PYRAMID, par, 0

    @property
    def type(self) -> str:
        return self._type

    @property
    def main_width(self) -> float:
        return self._main_width

    @property
    def center(self) -> pt.Tensor:
        return self._center
