"""
Device gather + weighted-sum interpolation (host wrapper around ``s3_interp_gather``).

``interpolate_data`` keeps the reference's signature (sparseSpatialSampling/export.py:446-468).
"""
import torch as pt

from . import _lib

_DT = {pt.float32: _lib.S3_F32, pt.float64: _lib.S3_F64}


def interp_gather(data: pt.Tensor, idx: pt.Tensor, weights: pt.Tensor, out: pt.Tensor = None,
                  out_row: pt.Tensor = None, out_dtype=None) -> pt.Tensor:
    """
    ``out[c] = sum_j weights[c, j] * data[idx[c, j]]`` on the device.

    :param data: CUDA tensor ``[N, ...]`` (fp32 or fp64), contiguous; trailing dims are flattened into one row
    :param idx: CUDA int32 ``[Nc, k]``
    :param weights: CUDA ``[Nc, k]``; fp32 for the fp32 fast path, fp64 for fp64 output
    :param out_row: optional CUDA int32 ``[Nc]``, destination row of each processed cell
    """
    _lib.require_cuda()
    lib = _lib.load()
    assert data.is_cuda and idx.is_cuda and weights.is_cuda
    assert idx.dtype == pt.int32 and idx.dim() == 2
    data = data.contiguous()
    n_src = data.size(0)
    row_len = data.numel() // max(n_src, 1)
    n_cells, k = idx.shape
    if out_dtype is None:
        out_dtype = pt.float32 if (data.dtype == pt.float32 and weights.dtype == pt.float32) else pt.float64
    want_w = pt.float32 if out_dtype == pt.float32 else pt.float64
    if weights.dtype != want_w:
        weights = weights.to(want_w)
    if out is None:
        out = pt.empty((n_cells,) + tuple(data.shape[1:]), dtype=out_dtype, device=data.device)
    assert out.is_contiguous() and out.dtype == out_dtype
    with pt.cuda.device(data.device):
        _lib.check(lib.s3_interp_gather(_lib.ptr(data), _DT[data.dtype], n_src, row_len, _lib.ptr(idx.contiguous()),
                                        _lib.ptr(weights.contiguous()), n_cells, k, _lib.ptr(out_row), _lib.ptr(out),
                                        _DT[out_dtype], _lib.stream_ptr()))
    return out


def interpolate_data(weights: pt.Tensor, idx_weights: pt.Tensor, data: pt.Tensor, chunk_size: int = 100000,
                     out_dtype=None) -> pt.Tensor:
    """
    Drop-in for the reference's ``interpolate_data(weights, idx_weights, data, chunk_size)``.

    Host tensors are moved to the current CUDA device, interpolated there and returned on the device of ``data``.
    ``chunk_size`` is accepted for signature compatibility; the kernel never materialises ``data[idx]`` so no
    chunking over cells is needed.
    """
    _lib.require_cuda()
    dev = pt.device("cuda", pt.cuda.current_device())
    src_device = data.device
    d = data.to(dev, non_blocking=True)
    i = idx_weights.to(dev).to(pt.int32)
    if out_dtype is None:
        out_dtype = weights.dtype if weights.dtype in (pt.float32, pt.float64) else pt.float64
        if d.dtype == pt.float64:
            out_dtype = pt.float64
    w = weights.to(dev)
    if d.dtype not in (pt.float32, pt.float64):
        d = d.to(pt.float64 if out_dtype == pt.float64 else pt.float32)
    out = interp_gather(d, i, w, out_dtype=out_dtype)
    return out if src_device.type == "cuda" else out.to(src_device)


class GroupTables:
    """
    Tables of the grouped interpolation kernel (``s3_interp_groups_build`` / ``s3_interp_grouped``): for every group of
    consecutive cells (processing order; group size ``s3_interp_group_size()`` = 4) the distinct source rows of the
    group, a membership mask and one weight per (row, cell). Built once per KNN cache; a warp then loads every distinct
    row once for all cells of its group instead of once per (cell, neighbour) reference.
    """

    def __init__(self, idx_sorted: pt.Tensor, w32_sorted: pt.Tensor):
        _lib.require_cuda()
        lib = _lib.load()
        dev = idx_sorted.device
        assert idx_sorted.dtype == pt.int32 and w32_sorted.dtype == pt.float32
        self.n_cells, self.k = idx_sorted.shape
        self.group = int(lib.s3_interp_group_size())
        self.n_groups = (self.n_cells + self.group - 1) // self.group
        cap = self.group * self.k
        self.cnt = pt.zeros((max(self.n_groups, 1),), dtype=pt.int32, device=dev)
        self.ent = pt.empty((max(self.n_groups, 1), cap, 2), dtype=pt.int32, device=dev)
        self.wts = pt.empty((max(self.n_groups, 1), cap, 4), dtype=pt.float32, device=dev)
        with pt.cuda.device(dev):
            _lib.check(lib.s3_interp_groups_build(_lib.ptr(idx_sorted.contiguous()), _lib.ptr(w32_sorted.contiguous()),
                                                  self.n_cells, self.k, _lib.ptr(self.cnt), _lib.ptr(self.ent),
                                                  _lib.ptr(self.wts), _lib.stream_ptr()))

    @property
    def rows_per_cell(self) -> float:
        """Distinct rows loaded per cell (k without grouping)."""
        return float(self.cnt.sum().item()) / max(self.n_cells, 1)

    def interpolate(self, data: pt.Tensor, out: pt.Tensor = None, out_row: pt.Tensor = None) -> pt.Tensor:
        lib = _lib.load()
        assert data.is_cuda and data.dtype == pt.float32
        data = data.contiguous()
        n_src = data.size(0)
        row_len = data.numel() // max(n_src, 1)
        if out is None:
            out = pt.empty((self.n_cells,) + tuple(data.shape[1:]), dtype=pt.float32, device=data.device)
        assert out.is_contiguous() and out.dtype == pt.float32
        with pt.cuda.device(data.device):
            _lib.check(lib.s3_interp_grouped(_lib.ptr(data), n_src, row_len, _lib.ptr(self.cnt), _lib.ptr(self.ent),
                                             _lib.ptr(self.wts), self.n_cells, self.k, _lib.ptr(out_row), _lib.ptr(out),
                                             _lib.S3_F32, _lib.stream_ptr()))
        return out


class StagedTiles:
    """
    Tile structures of the staged interpolation kernel (``s3_interp_tiles_build`` / ``s3_interp_staged``): for every
    tile of 32 consecutive cells (processing order) the ascending list of unique source rows and the position of
    every (cell, neighbour) reference in it. Built once per KNN cache.
    """
    TILE = 32

    def __init__(self, idx_sorted: pt.Tensor, w32_sorted: pt.Tensor):
        _lib.require_cuda()
        lib = _lib.load()
        dev = idx_sorted.device
        self.n_cells, self.k = idx_sorted.shape
        self.n_tiles = (self.n_cells + self.TILE - 1) // self.TILE
        cap = self.TILE * self.k
        self.rows = pt.empty((max(self.n_tiles, 1), cap), dtype=pt.int32, device=dev)
        self.nrows = pt.zeros((max(self.n_tiles, 1),), dtype=pt.int32, device=dev)
        self.lidx = pt.empty((max(self.n_tiles, 1), cap), dtype=pt.uint16, device=dev)
        self.w = pt.zeros((max(self.n_tiles, 1) * self.TILE, self.k), dtype=pt.float32, device=dev)
        self.w[:self.n_cells] = w32_sorted
        with pt.cuda.device(dev):
            _lib.check(lib.s3_interp_tiles_build(_lib.ptr(idx_sorted.contiguous()), self.n_cells, self.k,
                                                 _lib.ptr(self.rows), _lib.ptr(self.nrows), _lib.ptr(self.lidx),
                                                 _lib.stream_ptr()))
        self.max_rows = int(self.nrows.max().item()) if self.n_tiles else 1
        self.total_rows = int(self.nrows.sum().item()) if self.n_tiles else 0

    def interpolate(self, data: pt.Tensor, out: pt.Tensor = None, out_row: pt.Tensor = None,
                    chunk_cols: int = 256, pipelined: bool = False, stage_rows: int = 0, n_ctas: int = 0,
                    gather4: bool = True) -> pt.Tensor:
        lib = _lib.load()
        assert data.is_cuda and data.dtype == pt.float32
        data = data.contiguous()
        n_src = data.size(0)
        row_len = data.numel() // max(n_src, 1)
        if out is None:
            out = pt.empty((self.n_cells,) + tuple(data.shape[1:]), dtype=pt.float32, device=data.device)
        if pipelined:
            with pt.cuda.device(data.device):
                _lib.check(lib.s3_interp_pipelined(_lib.ptr(data), n_src, row_len, _lib.ptr(self.rows),
                                                   _lib.ptr(self.nrows), _lib.ptr(self.lidx), _lib.ptr(self.w),
                                                   self.n_cells, self.k, self.max_rows, int(chunk_cols),
                                                   int(stage_rows), int(n_ctas), int(gather4), _lib.ptr(out_row), _lib.ptr(out),
                                                   _lib.stream_ptr()))
            return out
        with pt.cuda.device(data.device):
            _lib.check(lib.s3_interp_staged(_lib.ptr(data), n_src, row_len, _lib.ptr(self.rows), _lib.ptr(self.nrows),
                                            _lib.ptr(self.lidx), _lib.ptr(self.w), self.n_cells, self.k, self.max_rows,
                                            int(chunk_cols), _lib.ptr(out_row), _lib.ptr(out), _lib.stream_ptr()))
        return out

