"""
Host-side cell topology (neighbour pointers, shared node ids) with the reference's history-dependent semantics
(sparseSpatialSampling/s_cube.py: ``Cell.nb``, ``Cell.node_idx``, ``_assign_neighbors``, ``_assign_indices``,
``_check_nb``, the neighbour reset of ``_remove_invalid_cells``, ``_resort_nodes_and_indices_of_grid``).
Thin ctypes wrapper around ``s3_topo_*`` (``csrc/topology.cu``, plain C++ on host arrays -- needs no device).
"""
import ctypes

import numpy as np

from . import _lib


def _i64(values) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(values, dtype=np.int64).reshape(-1))


class Topology:
    def __init__(self, n_dimensions: int, root_center, width: float, asynchronous: bool = False):
        """``asynchronous``: updates are queued and applied in order by a native worker thread (no Python involved);
        reads (``check_nb``, ``cell``, ``final``, the counts) wait for the queue; ``sync()`` reports a queued failure."""
        self._lib = _lib.load()
        self.n_dimensions = int(n_dimensions)
        self.n_children = 2 ** self.n_dimensions
        self.n_neighbours = 8 if self.n_dimensions == 2 else 26
        c = np.ascontiguousarray(np.asarray(root_center, dtype=np.float64).reshape(-1))
        h = ctypes.c_void_p()
        _lib.check(self._lib.s3_topo_create(self.n_dimensions, c.ctypes.data, float(width), int(bool(asynchronous)),
                                            ctypes.byref(h)))
        self._h = h
        self._nb_buf = np.zeros(26, dtype=np.int64)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._lib.s3_topo_free(h)
            self._h = None

    def __getstate__(self):
        raise TypeError("Topology holds a native handle and is not picklable (SamplingTree drops it after refine())")

    def sync(self) -> None:
        _lib.check(self._lib.s3_topo_sync(self._h))

    @property
    def n_cells(self) -> int:
        return int(self._lib.s3_topo_n_cells(self._h))

    @property
    def n_nodes(self) -> int:
        return int(self._lib.s3_topo_n_nodes(self._h))

    def refine(self, parents) -> None:
        """Children of ``parents`` (in that order) get the next indices; neighbours and node ids as the reference."""
        p = _i64(parents)
        _lib.check(self._lib.s3_topo_refine(self._h, p.ctypes.data, p.size))

    def refresh_siblings(self, cells) -> None:
        """``cell.parent.children = _assign_neighbors(cell.parent, children=cell.parent.children)`` per cell."""
        c = _i64(cells)
        _lib.check(self._lib.s3_topo_refresh(self._h, c.ctypes.data, c.size, 0))

    def refresh_children(self, parents) -> None:
        """Same for a list of parents (second pass of the uniform refinement)."""
        p = _i64(parents)
        _lib.check(self._lib.s3_topo_refresh(self._h, p.ctypes.data, p.size, 1))

    def mark_invalid(self, cells) -> None:
        c = _i64(cells)
        _lib.check(self._lib.s3_topo_mark_invalid(self._h, c.ctypes.data, c.size))

    def check_nb(self, cell: int) -> list:
        """Leaf neighbours with a lower level, in neighbour-slot order (``_check_nb``)."""
        n = int(self._lib.s3_topo_check_nb(self._h, int(cell), self._nb_buf.ctypes.data))
        if n < 0:
            raise _lib.S3Error(f"s3_topo_check_nb: bad cell {cell}")
        return self._nb_buf[:n].tolist()

    def cell(self, cell: int):
        """(nb int32 [8|26], node ids int32 [2^d], parent, children, level); children: first child, -1 leaf, -2 removed."""
        nb = np.zeros(self.n_neighbours, dtype=np.int32)
        nodes = np.zeros(self.n_children, dtype=np.int32)
        state = np.zeros(3, dtype=np.int32)
        _lib.check(self._lib.s3_topo_cell(self._h, int(cell), nb.ctypes.data, nodes.ctypes.data, state.ctypes.data))
        return nb, nodes, int(state[0]), int(state[1]), int(state[2])

    def final(self):
        """faces int32 [n_leaf, 2^d] (cell-list order), vertices fp64 [n_vertices, d], centres of all cells [n_cells, d]."""
        n_leaf, n_vert = ctypes.c_int64(0), ctypes.c_int64(0)
        _lib.check(self._lib.s3_topo_final(self._h, ctypes.byref(n_leaf), ctypes.byref(n_vert), None, None, None))
        faces = np.zeros((n_leaf.value, self.n_children), dtype=np.int32)
        vertices = np.zeros((n_vert.value, self.n_dimensions), dtype=np.float64)
        centers = np.zeros((self.n_cells, self.n_dimensions), dtype=np.float64)
        _lib.check(self._lib.s3_topo_final(self._h, ctypes.byref(n_leaf), ctypes.byref(n_vert), faces.ctypes.data,
                                           vertices.ctypes.data, centers.ctypes.data))
        return faces, vertices, centers
