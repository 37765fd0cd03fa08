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
        try:
            from . import build as _b
        except ImportError:  # imported as a top-level module
            import importlib.util
            spec = importlib.util.spec_from_file_location("s3_oracle_build", os.path.join(_HERE, "build.py"))
            _b = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(_b)
        try:
            _b.build()                      # no-op unless knn_oracle.c is newer than the library
        except Exception:
            if not os.path.exists(_SO):
                raise
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


# ---------------------------------------------------------------------------------- geometry masks
G_CUBE, G_SPHERE, G_CYLINDER, G_TRIANGLE, G_PRISM, G_TETRA, G_PYRAMID, G_STL, G_POLY2D = range(9)


def _flat(t):
    return [float(v) for v in np.asarray(t, dtype=np.float64).reshape(-1)]


def geometry_params(g):
    """
    (type id, fp64 parameters, n_extra) of a geometry object, read from the attributes the reference classes
    define (sparseSpatialSampling/geometry/*.py) -- works on reference objects and on the product's mirrors alike.
    """
    t = g.type
    if t == "cube":
        return G_CUBE, _flat(g._lower_bound) + _flat(g._upper_bound), 0
    if t == "sphere":
        return G_SPHERE, _flat(g._position) + [float(g._radius)], 0
    if t == "cylinder":
        cone = not isinstance(g._radius, (int, float))
        r0, r1 = (g._radius[0], g._radius[1]) if cone else (g._radius, g._radius)
        return G_CYLINDER, (_flat(g._position[0]) + _flat(g._axis) + [float(g._norm), float(r0), float(r1),
                                                                     1.0 if cone else 0.0]), 0
    if t == "triangle":
        return G_TRIANGLE, _flat(np.stack([np.asarray(p, dtype=np.float64) for p in g._points])), 0
    if t == "prism":
        dims = [int(v) for v in g._dim]
        tri = np.asarray(g._positions[0], dtype=np.float64)[:, dims]
        return G_PRISM, (_flat(g._positions[0][0]) + _flat(g._axis) + [float(g._norm), float(dims[0]), float(dims[1])]
                         + _flat(tri)), 0
    if t == "tetrahedron":
        return G_TETRA, _flat(g._positions) + _flat(np.asarray(g._normals, dtype=np.float64).T), 0
    if t == "pyramid":
        par = []
        for tet in g._tets:
            par += geometry_params(tet)[1]
        return G_PYRAMID, par, 0
    if t == "STL":
        tri = np.asarray(g._triangles, dtype=np.float64).reshape(-1, 9)
        return G_STL, _flat(g._lower_bound) + _flat(g._upper_bound) + [float(g._tolerance)] + _flat(tri), tri.shape[0]
    if t == "coord_2D":
        v = np.asarray(g._vertices, dtype=np.float64).reshape(-1, 2)
        return G_POLY2D, _flat(g._lower_bound) + _flat(g._upper_bound) + _flat(v), v.shape[0]
    raise ValueError(f"unknown geometry type {t}")


def points_inside(g, points: np.ndarray) -> np.ndarray:
    """Per-point inside mask of geometry ``g`` (the reference's mask_box / mask_sphere / _mask_* functions)."""
    lib = _load_c()
    if not hasattr(lib, "_s3o_pi_ready"):
        lib.s3o_points_inside.restype = None
        lib.s3o_points_inside.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64,
                                          ctypes.c_int, ctypes.c_void_p]
        lib._s3o_pi_ready = True
    type_id, par, n_extra = geometry_params(g)
    par = np.ascontiguousarray(par, dtype=np.float64)
    pts = np.ascontiguousarray(points, dtype=np.float64)
    out = np.empty(pts.shape[0], dtype=np.uint8)
    lib.s3o_points_inside(type_id, par.ctypes.data, n_extra, pts.ctypes.data, pts.shape[0], pts.shape[1],
                          out.ctypes.data)
    return out.astype(bool)


def apply_mask(mask: np.ndarray, keep_inside: bool, refine_geometry: bool) -> bool:
    """GeometryObject._apply_mask (geometry_base.py:40-76)."""
    if not refine_geometry:
        return bool(mask.all()) if not keep_inside else bool(not mask.any())
    return bool(mask.any()) if not keep_inside else bool(not mask.all())


def check_cell(g, nodes: np.ndarray, refine_geometry: bool = False) -> bool:
    return apply_mask(points_inside(g, nodes), g.keep_inside, refine_geometry)


# ------------------------------------------------------------------------------ refinement loop
DIRS_2D = np.array([[-1, -1], [-1, 1], [1, 1], [1, -1]], dtype=np.float64)
DIRS_3D = np.array([[-1, -1, 1], [-1, 1, 1], [1, 1, 1], [1, -1, 1], [-1, -1, -1], [-1, 1, -1], [1, 1, -1], [1, -1, -1]],
                   dtype=np.float64)


def gain_formula(level, n_dims, width, gain0, sdm):
    """numba fastmath _update_gain (s_cube.py:1840-1859), probed form: (((1/2^d) * c^d) * sdm) / gain0."""
    c = width / (2 ** level)
    q = c * c if n_dims == 2 else (c * c) * c
    return (((1 / (2 ** n_dims)) * q) * sdm) / gain0


def sum_delta_metric(m: np.ndarray, order: int) -> np.ndarray:
    """(|m0 - mj|).sum(dim=1) with torch's association order (s_cube.py:229); order 1 = four-lane (3-D, probed)."""
    a = np.abs(m[:, [0]] - m[:, 1:])
    if a.shape[1] == 8 and order == 1:
        return (((a[:, 0] + a[:, 4]) + (a[:, 1] + a[:, 5])) + (a[:, 2] + a[:, 6])) + (a[:, 3] + a[:, 7])
    s = a[:, 0].copy()
    for j in range(1, a.shape[1]):
        s = s + a[:, j]
    return s


class OracleTree:
    """
    CPU restatement of SamplingTree (s_cube.py:86-902, 1538-1584) on flat Python lists; control flow and the set
    operations that define the cell numbering follow the reference line by line.
    ``topology``: optional factory ``(d, root_center, width) -> object`` with the methods of
    ``oracle.topology_oracle.OracleTopology`` (or the product's ``Topology``); when given, the tree replays the
    reference's neighbour pointers / node ids (s_cube.py:904-1536) next to the refinement, uses them for the
    ``max_delta_level`` closure (s_cube.py:447-506, exact) and returns the reference's ``face_ids`` / ``all_nodes``.
    Without it ``max_delta_level`` is restated geometrically (the neighbour of a cell in a direction is the leaf covering
    the adjacent same-level lattice position), which equals the pointer version only while no pointer is stale.
    """

    def __init__(self, vertices, target, geometries, n_cells=None, uniform_level=5, min_metric=0.75,
                 n_cells_iter_start=None, n_cells_iter_end=None, relTol=1e-3, reach_at_least=0.75, pre_select=False,
                 sdm_order=1, max_delta_level=False, topology=None):
        self.X = np.ascontiguousarray(vertices, dtype=np.float64)
        self.y = np.ascontiguousarray(target, dtype=np.float64)
        self.geometries = geometries
        self.d = self.X.shape[1]
        self.k = 8 if self.d == 2 else 26
        self.nch = 2 ** self.d
        self.dirs = DIRS_2D if self.d == 2 else DIRS_3D
        self.n_cells_max, self.min_metric, self.min_level = n_cells, min_metric, uniform_level
        self.cpi_start = int(0.001 * self.X.shape[0]) if n_cells_iter_start is None else n_cells_iter_start
        if self.cpi_start <= 0:
            self.cpi_start = 1
        self.cpi_end = self.cpi_start if n_cells_iter_end is None else n_cells_iter_end
        self.cpi, self.cpi_last = self.cpi_start, 1e9
        self.relTol, self.reach_at_least, self.pre_select, self.sdm_order = relTol, reach_at_least, pre_select, sdm_order
        self.metric_log, self.n_cells_log = [], []
        self.max_delta_level = max_delta_level
        self.lattice, self.lookup = [], {}
        plane = [(-1, 0), (-1, 1), (0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1)]      # NB order, s_cube.py:22-26
        if self.d == 2:
            self.nb_dirs = plane
        else:
            self.nb_dirs = ([p + (0,) for p in plane] + [p + (-1,) for p in plane + [(0, 0)]] +
                            [p + (1,) for p in plane + [(0, 0)]])
        self.center, self.level, self.gain, self.metric, self.invalid = [], [], [], [], []
        self.leaf = set()
        self.iterations = 0
        # root (s_cube.py:338-397)
        mid = None
        for g in geometries:
            if g.keep_inside:
                self.width = float(g.main_width)
                mid = np.asarray(g.center, dtype=np.float64)
        c = np.repeat(mid[None, :], self.nch + 1, axis=0)
        c[1:] += self.dirs * 0.25 * self.width
        m = knn_predict(self.X, self.y, c, self.k)
        g0 = pow(self.width / 2, self.d) * sum([abs(m[0] - m[i]) for i in range(1, len(m))])
        if abs(g0 - 0) < 1e-6:
            g0 = 1.0
        self.gain0 = float(g0)
        self.center.append(c[0].copy()); self.level.append(0); self.gain.append(self.gain0)
        self.metric.append(float(m[0])); self.invalid.append(False)
        self.lattice.append((0,) * self.d); self.lookup[(0,) + (0,) * self.d] = 0
        self.leaf.add(0)
        self.target_norm = float(np.linalg.norm(self.y))
        self.topo = topology(self.d, c[0].copy(), self.width) if topology is not None else None

    def _refine_cells(self, to_refine):
        parents = list(to_refine)
        first = len(self.center)
        all_parents, all_children = set(), set()
        new_index = first
        for i in parents:
            off = self.dirs * 0.25 * self.width / (2 ** self.level[i])
            ch = self.center[i][None, :] + off
            for j in range(self.nch):
                self.center.append(ch[j]); self.level.append(self.level[i] + 1)
                self.gain.append(0.0); self.metric.append(0.0); self.invalid.append(False)
                lat = tuple(2 * self.lattice[i][a] + (1 if self.dirs[j][a] > 0 else 0) for a in range(self.d))
                self.lattice.append(lat)
                self.lookup[(self.level[i] + 1,) + lat] = len(self.center) - 1
            all_children.update(list(range(new_index, new_index + self.nch)))
            all_parents.add(i)
            new_index += self.nch
        self.leaf -= all_parents
        self.leaf.update(all_children)
        new = list(range(first, new_index))
        if self.topo is not None:
            self.topo.refine(parents)                         # _assign_neighbors + _assign_indices per parent
        self._update_gain(new)
        return new

    def _check_nb(self, c):
        if self.topo is not None:
            return self.topo.check_nb(c)
        lv, pos, out = self.level[c], self.lattice[c], []
        for dv in self.nb_dirs:
            q = tuple(pos[a] + dv[a] for a in range(self.d))
            if min(q) < 0 or max(q) >= (1 << lv):
                continue
            for up in range(lv + 1):
                n = self.lookup.get((lv - up,) + tuple(v >> up for v in q))
                if n is None:
                    continue
                if up > 0 and n in self.leaf:
                    out.append(n)
                break
        return out

    def _check_constraint(self, viol):
        go = True if viol else False
        while go:
            tmp = set()
            for c in viol:
                if self.topo is not None:
                    self.topo.refresh_siblings([c])           # s_cube.py:489-490
                tmp.update(self._check_nb(c))
            if not tmp or tmp.issubset(viol):
                go = False
            else:
                viol.update(tmp)
        return viol

    def _update_gain(self, cells):
        if not cells:
            return
        n = len(cells)
        q = np.empty((n, self.nch + 1, self.d))
        for t, i in enumerate(cells):
            q[t, 0] = self.center[i]
            q[t, 1:] = self.center[i][None, :] + self.dirs * 0.25 * self.width / (2 ** self.level[i])
        m = knn_predict(self.X, self.y, q.reshape(-1, self.d), self.k).reshape(n, self.nch + 1)
        sdm = sum_delta_metric(m, self.sdm_order)
        for t, i in enumerate(cells):
            self.gain[i] = gain_formula(self.level[i], self.d, self.width, self.gain0, float(sdm[t]))
            self.metric[i] = float(m[t, 0])

    def _nodes(self, i):
        return self.center[i][None, :] + self.dirs * 0.5 * self.width / (2 ** self.level[i])

    def _remove_invalid_cells(self, refined, refine_geometry=False, geometry_no=None):
        if self.pre_select:
            return None
        geoms = self.geometries if geometry_no is None else [self.geometries[geometry_no]]
        result = []
        for c in refined:
            nodes = self._nodes(c)
            hit = None
            for g in geoms:
                if check_cell(g, nodes, refine_geometry):
                    hit = c
                    break
            result.append(hit)
        idx = set(filter(None, result))
        if idx == set():
            return None
        if refine_geometry:
            return idx
        for c in idx:
            self.invalid[c] = True
            self.gain[c] = 0
        if self.topo is not None:
            self.topo.mark_invalid(list(idx))                 # s_cube.py:721-731
        self.leaf -= idx
        return None

    def _captured(self):
        cur = np.array([self.metric[c] for c in self.leaf])
        self.metric_log.append(float(np.linalg.norm(cur) / self.target_norm))

    def _continue(self):
        if self.n_cells_max is None:
            if len(self.metric_log) > 1 and self.metric_log[-1] / self.min_metric >= self.reach_at_least:
                return self.metric_log[-1] < self.min_metric and abs(self.metric_log[-1] - self.metric_log[-2]) > self.relTol
        else:
            if len(self.leaf) / self.n_cells_max >= self.reach_at_least:
                rel = abs(self.cpi / self.n_cells_max - self.cpi_last / self.n_cells_max)
                return len(self.leaf) < self.n_cells_max and rel > self.relTol
        return True

    def _update_cpi(self):
        if self.n_cells_max is None:
            dx, cx = self.min_metric - self.metric_log[0], self.metric_log[-1]
        else:
            dx, cx = self.n_cells_max - self.n_after_uniform, len(self.center)
        new = self.cpi_start - ((self.cpi_start - self.cpi_end) / dx) * cx
        self.cpi_last = self.cpi
        self.cpi = int(new) if new > 1 else 1

    def refine(self):
        import heapq
        for _ in range(self.min_level):
            parents = list(self.leaf)
            new = self._refine_cells(self.leaf)
            if self.topo is not None:
                self.topo.refresh_children(parents)           # second pass, s_cube.py:547-549
            self._remove_invalid_cells({c for c in new})
        self.n_after_uniform = len(self.leaf)
        if self.n_cells_max is None:
            self._captured()
        self.n_cells_log.append(len(self.leaf))
        self.selected_log = []
        while self._continue():
            if len(self.metric_log) >= 2:
                self._update_cpi()
            srt = heapq.nlargest(min(self.cpi, len(self.center)), self.leaf, key=lambda i: (self.gain[i], -i))
            self.selected_log.append(list(srt))
            to_refine = set()
            for i in srt:
                to_refine.add(i)
                if self.topo is not None:
                    self.topo.refresh_siblings([i])           # s_cube.py:611
                if self.max_delta_level:
                    to_refine.update(self._check_constraint(set(self._check_nb(i))))
            self._remove_invalid_cells({c for c in self._refine_cells(to_refine)})
            if self.n_cells_max is None:
                self._captured()
            self.iterations += 1
            self.n_cells_log.append(len(self.leaf))
        if self.n_cells_max is not None:
            self._captured()
        # geometry refinement (s_cube.py:774-863)
        for gi, g in enumerate(self.geometries):
            if not g.refine:
                continue
            found = self._remove_invalid_cells(self.leaf, True, gi)
            if found is None:
                break
            cells = set(found)
            lo = min(self.level[c] for c in cells)
            hi = max(self.level[c] for c in cells) if g.min_refinement_level is None else g.min_refinement_level
            while hi > lo:
                to_refine, checked = set(), set()
                for i in cells:
                    if i in checked:
                        continue
                    if self.level[i] < hi:
                        to_refine.add(i)
                        if self.topo is not None:
                            self.topo.refresh_siblings([i])   # s_cube.py:826
                    if self.max_delta_level:
                        more = set(self._check_nb(i))
                        more.update(self._check_constraint(more))
                        to_refine.update(more)
                        checked.update(more)
                idx_new = {c for c in self._refine_cells(to_refine)}
                self._remove_invalid_cells(idx_new, geometry_no=gi)
                found = self._remove_invalid_cells({i for i in idx_new if not self.invalid[i]}, True, gi)
                if found is None:
                    break
                cells = set(found)
                lo += 1
        order = list(self.leaf)
        self.all_centers = np.stack([self.center[c] for c in order])
        self.all_levels = np.array([self.level[c] for c in order], dtype=np.int64)[:, None]
        self.leaf_order = order
        if self.topo is not None:
            self.face_ids, self.all_nodes, _ = self.topo.final()
        return self


def interpolate_torch(weights, idx, data, chunk_size: int = 100000):
    """
    The reference's CPU evaluation strategy for interpolate_data (export.py:446-468) with torch CPU operators
    (all intra-op threads): gather the k source rows of a chunk of cells, multiply by the weights, reduce over k.
    Used as the timed CPU baseline (bench.py) -- same temporaries, same threading model as the reference.
    """
    import torch as pt
    nc, k = idx.shape
    out = pt.empty((nc, data.shape[1], data.shape[2]), dtype=weights.dtype)
    for s in range(0, nc, chunk_size):
        e = min(s + chunk_size, nc)
        rows = data.index_select(0, idx[s:e].reshape(-1)).reshape(e - s, k, data.shape[1], data.shape[2])
        out[s:e] = (rows * weights[s:e].reshape(e - s, k, 1, 1)).sum(dim=1)
    return out


# ---------------------------------------------------------------------------------------------- weighted SVD
def cell_area(size_initial_cell: float, levels: np.ndarray, n_dims: int) -> np.ndarray:
    """Dataloader._compute_cell_area (data.py:240-247): (size_initial_cell / 2^level)^d per cell."""
    return np.power(size_initial_cell / np.power(2.0, np.asarray(levels, dtype=np.float64)), n_dims).squeeze()


def compute_svd(data_matrix: np.ndarray, area: np.ndarray, rank: int = None):
    """compute_svd (utils.py:302-346) in the reference's order of operations and dtype (fp32 in, fp32 SVD):
    subtract the temporal mean (:322), scale rows by sqrt(cell_area) (:326 / :335), stack vector components to
    (N_cells * D, T) rows (:337-338), thin SVD truncated to `rank` (:328 / :341 -- flowtorch.analysis.SVD is
    torch.linalg.svd(full_matrices=False) followed by truncation; flowtorch, branch `aweiner`, is not vendored, so the
    rank=None rule is not restated and an explicit rank is required), un-scale U (:330 / :344-346).
    Returns (s [r], U [N_cells, r] | [N_cells, D, r], V [T, r]). PARITY UNPINNED by reference tests (none cover it);
    pinned against numpy's fp64 SVD in tests/test_oracle.py."""
    assert rank is not None, "the oracle needs an explicit rank"
    a = np.array(data_matrix, dtype=np.float32, copy=True)
    w = np.sqrt(np.asarray(area, dtype=np.float32))
    a -= a.mean(axis=-1, keepdims=True, dtype=np.float32)
    if a.ndim == 2:
        a *= w[:, None]
        u, s, vh = np.linalg.svd(a, full_matrices=False)
        r = min(rank, s.shape[0])
        return s[:r], u[:, :r] / w[:, None], vh.T[:, :r]
    a *= w[:, None, None]
    shape = a.shape
    u, s, vh = np.linalg.svd(a.reshape(shape[0] * shape[1], shape[2]), full_matrices=False)
    r = min(rank, s.shape[0])
    return s[:r], u[:, :r].reshape(shape[0], shape[1], r) / w[:, None, None], vh.T[:, :r]


def weighted_gram(data_matrix: np.ndarray, area: np.ndarray) -> np.ndarray:
    """fp64 Gram matrix of the centred, sqrt(area)-weighted rows: the quantity the device contraction computes."""
    a = np.asarray(data_matrix, dtype=np.float64)
    if a.ndim == 3:
        area = np.repeat(np.asarray(area, dtype=np.float64), a.shape[1])
        a = a.reshape(a.shape[0] * a.shape[1], a.shape[2])
    mean32 = np.asarray(data_matrix, dtype=np.float32).reshape(a.shape).mean(axis=1, dtype=np.float64).astype(np.float32)
    b = (a.astype(np.float32) - mean32[:, None]).astype(np.float64) * np.sqrt(np.asarray(area, dtype=np.float32)).astype(np.float64)[:, None]
    return b.T @ b
