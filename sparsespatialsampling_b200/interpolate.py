"""
Device gather + weighted-sum interpolation (host wrapper around ``s3_interp_gather_strided``).

``interpolate_data`` keeps the reference's signature (sparseSpatialSampling/export.py:446-468).

Layout in HBM. The kernel reads the reference's ``[N, D, T]`` snapshot batches (T contiguous) through explicit strides,
so any view whose last dimension is contiguous is interpolated in place. It is built for rows whose pitch is a multiple
of 128 bytes: ``alloc_snapshots`` / ``to_pitched`` give such tensors (``[N, D, T]`` views of a ``[N, D, Tp]``
allocation, ``Tp`` = T rounded up to 32 fp32 values), the streamed host path (``KnnTables.interpolate_host``) stages
its windows that way by construction.
"""
import torch as pt

from . import _lib

_DT = {pt.float32: _lib.S3_F32, pt.float64: _lib.S3_F64}
LINE_BYTES = 128


def pitched_columns(n_cols: int, dtype=pt.float32) -> int:
    """Columns per row so that the row pitch is a multiple of a 128-byte cache line."""
    per_line = LINE_BYTES // pt.empty((), dtype=dtype).element_size()
    return (int(n_cols) + per_line - 1) // per_line * per_line


def alloc_snapshots(n_rows: int, n_comp: int, n_cols: int, dtype=pt.float32, device=None, zero: bool = False) -> pt.Tensor:
    """``[n_rows, n_comp, n_cols]`` view with 128-byte aligned component rows (the layout the gather kernel is built for)."""
    tp = pitched_columns(n_cols, dtype)
    make = pt.zeros if zero else pt.empty
    return make((n_rows, n_comp, tp), dtype=dtype, device=device)[:, :, :n_cols]


def to_pitched(data: pt.Tensor) -> pt.Tensor:
    """Copy of a ``[N, D, T]`` / ``[N, T]`` device tensor in the pitched layout (no copy if it already is)."""
    if data.dim() == 2:
        return to_pitched(data.unsqueeze(1)).squeeze(1)
    if is_pitched(data):
        return data
    out = alloc_snapshots(data.size(0), data.size(1), data.size(2), data.dtype, data.device)
    out.copy_(data)
    return out


def is_pitched(data: pt.Tensor) -> bool:
    es = data.element_size()
    return (data.stride(-1) == 1 and all((s * es) % LINE_BYTES == 0 for s in data.stride()[:-1])
            and data.data_ptr() % 16 == 0)


def _rows_view(t: pt.Tensor):
    """(tensor, n_comp, n_cols, row_stride, comp_stride) of a ``[N, ...]`` tensor whose last dimension is contiguous;
    anything the kernel cannot address through two strides is made dense first."""
    if t.dim() == 1:
        t = t.unsqueeze(1)
    if t.dim() == 2:
        if t.stride(1) != 1 and t.size(1) > 1:
            t = t.contiguous()
        return t, 1, t.size(1), (t.stride(0) if t.size(0) > 1 else t.size(1)), t.size(1)
    if t.dim() > 3 or (t.stride(-1) != 1 and t.size(-1) > 1):
        t = t.contiguous().reshape(t.size(0), -1, t.size(-1)) if t.dim() > 3 else t.contiguous()
    n, d, c = t.shape
    comp_stride = t.stride(1) if d > 1 else c
    row_stride = t.stride(0) if n > 1 else max((d - 1) * comp_stride + c, 1)
    if comp_stride < c or row_stride < (d - 1) * comp_stride + c:          # overlapping / expanded views
        t = t.contiguous()
        comp_stride, row_stride = c, d * c
    return t, d, c, row_stride, comp_stride


def interp_gather(data: pt.Tensor, idx: pt.Tensor, weights: pt.Tensor, out: pt.Tensor = None,
                  out_row: pt.Tensor = None, out_dtype=None) -> pt.Tensor:
    """
    ``out[c] = sum_j weights[c, j] * data[idx[c, j]]`` on the device.

    :param data: CUDA tensor ``[N, T]`` / ``[N, D, T]`` (fp32 or fp64); any row / component stride, T contiguous
    :param idx: CUDA int32 ``[Nc, k]``
    :param weights: CUDA ``[Nc, k]``; fp32 for the fp32 fast path, fp64 for fp64 output
    :param out: optional result tensor ``[Nc, ...]`` of the same trailing shape (may be a pitched view)
    :param out_row: optional CUDA int32 ``[Nc]``, destination row of each processed cell
    """
    _lib.require_cuda()
    lib = _lib.load()
    assert data.is_cuda and idx.is_cuda and weights.is_cuda
    assert idx.dtype == pt.int32 and idx.dim() == 2
    n_src = data.size(0)
    n_cells, k = idx.shape
    if out_dtype is None:
        out_dtype = pt.float32 if (data.dtype == pt.float32 and weights.dtype == pt.float32) else pt.float64
    want_w = pt.float32 if out_dtype == pt.float32 else pt.float64
    if weights.dtype != want_w:
        weights = weights.to(want_w)
    shape_out = (n_cells,) + tuple(data.shape[1:])
    d, n_comp, n_cols, row_stride, comp_stride = _rows_view(data)
    if out is None:
        out = pt.empty(shape_out, dtype=out_dtype, device=data.device)
    assert out.dtype == out_dtype and tuple(out.shape) == shape_out, (out.dtype, tuple(out.shape), shape_out)
    o, o_comp, o_cols, o_row_stride, o_comp_stride = _rows_view(out)
    assert o.data_ptr() == out.data_ptr(), "the result tensor must have a contiguous last dimension and simple strides"
    with pt.cuda.device(data.device):
        _lib.check(lib.s3_interp_gather_strided(
            _lib.ptr(d, strided=True), _DT[data.dtype], n_src, n_comp, n_cols, row_stride, comp_stride,
            _lib.ptr(idx.contiguous()), _lib.ptr(weights.contiguous()), n_cells, k, _lib.ptr(out_row),
            _lib.ptr(o, strided=True), _DT[out_dtype], o_row_stride, o_comp_stride, _lib.stream_ptr()))
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

