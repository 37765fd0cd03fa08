"""GPU parity: s3_interp_gather against the CPU oracle of interpolate_data (export.py:446-468)."""
import numpy as np
import pytest
import torch as pt

from oracle import s3_oracle as orc

pytestmark = pytest.mark.gpu

# tolerance stated by BASELINE.json north_star: fp32 relative 1e-5 (relative to the largest gathered magnitude)
RTOL_F32 = 1e-5


def _case(N, Nc, k, D, T, seed):
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((N, D, T)).astype(np.float32)
    idx = rng.integers(0, N, (Nc, k)).astype(np.int32)
    w = rng.random((Nc, k))
    w /= w.sum(1, keepdims=True)
    return data, idx, w


@pytest.mark.parametrize("N,Nc,k,D,T", [
    (5000, 1777, 8, 1, 1000), (5000, 1000, 8, 2, 500), (3000, 515, 26, 3, 64), (2000, 100, 8, 1, 301),
    (2000, 33, 26, 1, 7), (100, 1, 8, 1, 4), (4000, 2049, 5, 1, 128),
])
def test_fp32_path_within_tolerance(cuda, N, Nc, k, D, T):
    from sparsespatialsampling_b200.interpolate import interp_gather
    data, idx, w = _case(N, Nc, k, D, T, N + Nc)
    ref = orc.interpolate(w, idx.astype(np.int64), data)
    out = interp_gather(pt.from_numpy(data).cuda(), pt.from_numpy(idx).cuda(), pt.from_numpy(w).float().cuda())
    assert out.dtype == pt.float32 and tuple(out.shape) == (Nc, D, T)
    scale = np.abs(data[idx]).max(axis=1)                        # max_k |data[idx[c,k]]|
    err = np.abs(out.cpu().numpy().astype(np.float64) - ref)
    assert (err <= RTOL_F32 * np.maximum(scale, 1e-30)).all(), err.max()


@pytest.mark.parametrize("N,Nc,k,D,T", [(3000, 700, 8, 2, 100), (3000, 300, 26, 1, 33)])
def test_fp64_path_bit_exact_vs_oracle(cuda, N, Nc, k, D, T):
    from sparsespatialsampling_b200.interpolate import interp_gather
    data, idx, w = _case(N, Nc, k, D, T, 3)
    ref = orc.interpolate(w, idx.astype(np.int64), data)
    out = interp_gather(pt.from_numpy(data).cuda(), pt.from_numpy(idx).cuda(), pt.from_numpy(w).cuda(),
                        out_dtype=pt.float64)
    assert np.array_equal(out.cpu().numpy(), ref)
    data64 = data.astype(np.float64)
    out64 = interp_gather(pt.from_numpy(data64).cuda(), pt.from_numpy(idx).cuda(), pt.from_numpy(w).cuda())
    assert np.array_equal(out64.cpu().numpy(), orc.interpolate(w, idx.astype(np.int64), data64))


def test_out_row_permutation_and_linearity(cuda):
    from sparsespatialsampling_b200.interpolate import interp_gather
    data, idx, w = _case(4000, 1234, 8, 1, 256, 9)
    d = pt.from_numpy(data).cuda()
    i = pt.from_numpy(idx).cuda()
    wt = pt.from_numpy(w).float().cuda()
    base = interp_gather(d, i, wt)
    perm = pt.randperm(1234, device="cuda")
    out = pt.zeros_like(base)
    interp_gather(d, i[perm].contiguous(), wt[perm].contiguous(), out=out, out_row=perm.to(pt.int32))
    assert pt.equal(out, base)
    # linearity in the data: I(a*x) == a*I(x) for a power of two (exact in fp32)
    assert pt.equal(interp_gather(d * 4.0, i, wt), base * 4.0)
    # constant field is reproduced (weights sum to one)
    ones = pt.ones_like(d)
    assert pt.allclose(interp_gather(ones, i, wt), pt.ones_like(base), rtol=0, atol=5e-7)


def test_reference_signature_interpolate_data(cuda):
    # interpolate_data(weights, idx_weights, data, chunk_size) with host tensors, fp64 weights -> fp64 result
    from sparsespatialsampling_b200.interpolate import interpolate_data
    data, idx, w = _case(2000, 500, 8, 2, 50, 21)
    out = interpolate_data(pt.from_numpy(w), pt.from_numpy(idx.astype(np.int64)), pt.from_numpy(data), 100)
    assert out.dtype == pt.float64 and not out.is_cuda
    assert np.array_equal(out.numpy(), orc.interpolate(w, idx.astype(np.int64), data))


@pytest.mark.parametrize("D,T,chunk", [(1, 1000, 256), (2, 300, 128), (3, 77, 32), (1, 64, None)])
def test_streamed_host_path_equals_resident_path(cuda, D, T, chunk):
    """Pipelined host -> device -> host export over windows of the time axis (pitched copies) == one resident launch."""
    from sparsespatialsampling_b200.export import KnnTables
    from sparsespatialsampling_b200.knn import KnnIndex
    rng = np.random.default_rng(D * T)
    x = pt.from_numpy(rng.random((6000, 2)))
    q = pt.from_numpy(rng.random((2500, 2)))
    data = pt.from_numpy(rng.standard_normal((6000, D, T)).astype(np.float32))
    tables = KnnTables(KnnIndex(x.cuda()), q.cuda(), 8)
    want = tables.interpolate(data.cuda(), pt.float32).cpu()
    got = tables.interpolate_host(data.pin_memory(), chunk_snapshots=chunk)
    assert not got.is_cuda and got.is_pinned() and pt.equal(got, want)
    dma = tables.interpolate_host(data.pin_memory(), chunk_snapshots=chunk, gather=False)   # pitched DMA of all rows
    assert pt.equal(dma, want)
    zc = tables.interpolate_host(data.pin_memory(), chunk_snapshots=chunk, gather=True)     # PCIe gather of referenced rows
    assert pt.equal(zc, want)
    # a grid that references few source points takes the gather path by itself
    few = KnnTables(KnnIndex(x.cuda()), q[:40].cuda(), 8)
    assert few._compact()[0].numel() < 0.6 * x.size(0)
    assert pt.equal(few.interpolate_host(data.pin_memory(), chunk_snapshots=chunk),
                    few.interpolate(data.cuda(), pt.float32).cpu())
    again = tables.interpolate_host(data, out=got, chunk_snapshots=chunk)       # pageable input, buffers re-used
    assert again.data_ptr() == got.data_ptr() and pt.equal(again, want)


def test_export_data_streams_host_batches(cuda, tmp_path):
    from sparsespatialsampling_b200.export import ExportData
    rng = np.random.default_rng(5)
    x = pt.from_numpy(rng.random((5000, 2)))

    class _Grid:
        pass
    g = _Grid()
    g.n_dimensions, g.faces, g.vertices, g.levels = 2, None, None, None
    g.centers, g.metric, g.size_initial_cell = pt.from_numpy(rng.random((1500, 2))), pt.from_numpy(rng.random(5000)), 1.0
    g.save_path, g.save_name, g.grid_name = str(tmp_path), "c", "grid"
    data = pt.from_numpy(rng.standard_normal((5000, 2, 512)).astype(np.float32))
    outs = []
    for stream in (True, False):
        exp = ExportData(g, write_times=[str(i) for i in range(512)], write_files=False, stream_host=stream)
        exp._stream_min_elements = 0
        exp.export(x, data, "U")
        outs.append(exp.interpolated_fields.centers)
    assert not outs[0].is_cuda and outs[1].is_cuda and pt.equal(outs[0], outs[1].cpu())


def _local_case(N, Nc, k, D, T, seed):
    """Neighbouring cells share neighbours (like Morton-ordered cells of a real grid): idx from a sliding window."""
    rng = np.random.default_rng(seed)
    data = rng.standard_normal((N, D, T)).astype(np.float32)
    base = np.sort(rng.integers(0, max(N - 3 * k, 1), Nc))
    idx = np.stack([rng.permutation(3 * k)[:k] + b for b in base]).astype(np.int32)   # distinct per cell
    w = rng.random((Nc, k))
    w /= w.sum(1, keepdims=True)
    return data, idx, w


@pytest.mark.parametrize("N,Nc,k,D,T", [
    (5000, 1777, 8, 1, 1000), (5000, 1002, 8, 2, 500), (3000, 515, 26, 3, 64), (2000, 101, 8, 1, 301),
    (2000, 33, 26, 1, 7), (100, 1, 8, 1, 4), (4000, 2049, 5, 1, 125), (3000, 64, 64, 2, 250),
])
def test_pitched_layout_equals_dense_layout(cuda, N, Nc, k, D, T):
    """Rows padded to a multiple of 128 bytes (the layout the kernel is built for), as source, as result and both:
    bit-identical to the dense launch -- including T % 4 != 0, where the dense layout has to take the scalar path."""
    from sparsespatialsampling_b200.interpolate import interp_gather, alloc_snapshots, to_pitched, is_pitched
    data, idx, w = _local_case(N, Nc, k, D, T, N + Nc + k)
    d_dense = pt.from_numpy(data).cuda()
    d_idx, d_w = pt.from_numpy(idx).cuda(), pt.from_numpy(w).float().cuda()
    want = interp_gather(d_dense, d_idx, d_w)
    d_pit = to_pitched(d_dense)
    assert is_pitched(d_pit) and (d_pit.stride(1) * 4) % 128 == 0 and pt.equal(d_pit, d_dense)
    assert pt.equal(interp_gather(d_pit, d_idx, d_w), want)
    out = alloc_snapshots(Nc, D, T, device="cuda", zero=True)
    interp_gather(d_pit, d_idx, d_w, out=out)
    assert pt.equal(out, want)
    assert float(out._base[:, :, T:].abs().sum()) == 0.0                  # nothing written into the padding
    out2 = alloc_snapshots(Nc, D, T, device="cuda")
    interp_gather(d_dense, d_idx, d_w, out=out2)
    assert pt.equal(out2, want)
    # reference dtype on the pitched layout: bit-exact against the oracle
    got64 = interp_gather(d_pit, d_idx, pt.from_numpy(w).cuda(), out_dtype=pt.float64)
    assert np.array_equal(got64.cpu().numpy(), orc.interpolate(w, idx.astype(np.int64), data))


@pytest.mark.parametrize("offset", [4, 8, 20, 28, 3])
def test_time_window_views_with_a_common_line_offset(cuda, offset):
    """A window ``data[:, :, t0:t1]`` of a pitched batch: every row starts `t0` columns into a 128-byte line; the kernel
    shortens its first step so that the following ones are line aligned. Odd offsets take the scalar path."""
    from sparsespatialsampling_b200.interpolate import interp_gather, alloc_snapshots
    data, idx, w = _local_case(3000, 700, 8, 2, 600, offset)
    full = alloc_snapshots(3000, 2, 600, device="cuda")
    full.copy_(pt.from_numpy(data))
    d_idx, d_w = pt.from_numpy(idx).cuda(), pt.from_numpy(w).float().cuda()
    for t1 in (600, 517, offset + 1):
        view = full[:, :, offset:t1]
        want = interp_gather(view.contiguous(), d_idx, d_w)
        assert pt.equal(interp_gather(view, d_idx, d_w), want), t1
        out = alloc_snapshots(700, 2, 600, device="cuda", zero=True)
        interp_gather(view, d_idx, d_w, out=out[:, :, offset:t1])
        assert pt.equal(out[:, :, offset:t1], want) and float(out[:, :, :offset].abs().sum()) == 0.0
        assert float(out[:, :, t1:].abs().sum()) == 0.0


@pytest.mark.parametrize("tune", ["1=4", "1=16", "2=1", "2=2", "3=0", "3=1", "4=256", "4=128,2=2", "6=1", "6=1,2=2",
                                  "6=0", "7=4", "7=8", "7=1", "7=4,3=1,2=2", "6=1,7=4", "9=0", "9=0,6=0,2=1", "16=0", "16=1,3=0",
                                  "16=1,2=2"])
def test_launch_variants_are_bit_identical(cuda, tune):
    """The launch knobs of the A/B harness (warps per CTA, column vectors per step, shared-memory pairs, column
    windows, 256-bit loads, neighbour-loop batching) change scheduling only: same sums in the same order."""
    from sparsespatialsampling_b200 import _lib
    from sparsespatialsampling_b200.interpolate import interp_gather, to_pitched
    cases = [_local_case(3000, 900, 8, 1, 1000, 1), _local_case(3000, 300, 26, 2, 333, 2)]
    defaults = {1: 8, 2: 0, 3: -1, 4: 0, 6: -1, 7: 0, 9: 1, 16: 1}
    _lib.tune(8, 0)                                   # reference results: the warp-per-cell kernel
    want = []
    for data, idx, w in cases:
        want.append(interp_gather(to_pitched(pt.from_numpy(data).cuda()), pt.from_numpy(idx).cuda(),
                                  pt.from_numpy(w).float().cuda()))
    try:
        for kv in tune.split(","):
            key, value = kv.split("=")
            _lib.tune(int(key), int(value))
        for (data, idx, w), ref in zip(cases, want):
            for layout in (to_pitched, lambda x: x):
                got = interp_gather(layout(pt.from_numpy(data).cuda()), pt.from_numpy(idx).cuda(),
                                    pt.from_numpy(w).float().cuda())
                assert pt.equal(got, ref)
            want64 = orc.interpolate(w, idx.astype(np.int64), data)
            for layout in (to_pitched, lambda x: x):           # pitched: the 256-bit path of the fp64-result mode
                got64 = interp_gather(layout(pt.from_numpy(data).cuda()), pt.from_numpy(idx).cuda(),
                                      pt.from_numpy(w).cuda(), out_dtype=pt.float64)
                assert np.array_equal(got64.cpu().numpy(), want64)
            d64 = to_pitched(pt.from_numpy(data).double().cuda())
            got64 = interp_gather(d64, pt.from_numpy(idx).cuda(), pt.from_numpy(w).cuda(), out_dtype=pt.float64)
            assert np.array_equal(got64.cpu().numpy(), orc.interpolate(w, idx.astype(np.int64), data.astype(np.float64)))
    finally:
        for key, value in defaults.items():
            _lib.tune(key, value)
        _lib.tune(8, PARTWARP_DEFAULT)


PARTWARP_DEFAULT = -1


@pytest.mark.parametrize("T", [125, 250, 61, 8, 100, 333, 1000])
@pytest.mark.parametrize("k,D", [(8, 1), (8, 2), (26, 3), (5, 1), (40, 1)])
def test_partwarp_kernel_is_bit_identical(cuda, T, k, D):
    """interp_partwarp_kernel (short rows of a sharded export: persistent part-warps, tables of the next cell fetched
    ahead, whole 256-bit vectors loaded across the row end) against the warp-per-cell kernel: same bits for every
    lanes-per-cell / load batch / CTA size, nothing written behind the last column of a row."""
    from sparsespatialsampling_b200 import _lib
    from sparsespatialsampling_b200.interpolate import interp_gather, to_pitched, alloc_snapshots
    data, idx, w = _local_case(2500, 1111, k, D, T, 5)
    d = to_pitched(pt.from_numpy(data).cuda())
    i_d, w_d = pt.from_numpy(idx).cuda(), pt.from_numpy(w).float().cuda()
    perm = pt.randperm(1111, generator=pt.Generator().manual_seed(0)).to(pt.int32).cuda()
    knobs = ((8, PARTWARP_DEFAULT), (12, 0), (13, 0), (7, 0), (1, 8))
    try:
        _lib.tune(8, 0)
        want = interp_gather(d, i_d, w_d)
        want_perm = interp_gather(d, i_d, w_d, out_row=perm)
        for tune in ("8=1", "8=1,7=8", "8=1,12=4", "8=1,12=3", "8=1,13=32", "8=1,13=16,7=8", "8=1,1=4", "8=1,1=3,13=16", "8=-1"):
            for key, value in knobs:
                _lib.tune(key, value)
            for kv in tune.split(","):
                key, value = kv.split("=")
                _lib.tune(int(key), int(value))
            out = alloc_snapshots(1111, D, T, device=d.device)
            out._base.fill_(-7.0)
            got = interp_gather(d, i_d, w_d, out=out)
            assert pt.equal(got, want), tune
            if out._base.size(-1) > T:
                assert bool((out._base[..., T:] == -7.0).all()), tune          # row padding untouched
            # dense result (rows aligned to 32 / 16 / 8 / 4 bytes depending on T): 256-bit loads, narrower stores
            assert pt.equal(interp_gather(d, i_d, w_d, out_row=perm), want_perm), tune
            assert pt.equal(interp_gather(d, i_d, w_d), want), tune
    finally:
        for key, value in knobs:
            _lib.tune(key, value)


def test_streamed_export_hands_out_fresh_results_and_releases_the_input(cuda, tmp_path):
    """export() of a pinned host batch must not return while the copy engine still reads the caller's tensor,
    and a result kept from batch i must survive batch i+1 (the reference returns fresh tensors)."""
    from sparsespatialsampling_b200.export import ExportData
    rng = np.random.default_rng(5)
    x = pt.from_numpy(rng.random((5000, 2)))

    class _Grid:
        pass
    g = _Grid()
    g.n_dimensions, g.faces, g.vertices, g.levels = 2, None, None, None
    g.centers, g.metric, g.size_initial_cell = pt.from_numpy(rng.random((1500, 2))), pt.from_numpy(rng.random(5000)), 1.0
    g.save_path, g.save_name, g.grid_name = str(tmp_path), "c", "grid"
    a = pt.from_numpy(rng.standard_normal((5000, 1, 512)).astype(np.float32))
    b = pt.from_numpy(rng.standard_normal((5000, 1, 512)).astype(np.float32))
    exp = ExportData(g, write_times=[str(i) for i in range(1024)], write_files=False)
    exp._stream_min_elements = 0
    buf = a.clone().pin_memory()
    exp.export(x, buf, "p", n_snapshots_total=1024)
    first = exp._last_fields.centers                     # kept across the next batch
    buf.copy_(b)                                         # refill the SAME pinned buffer, as a batch loop does
    exp.export(x, buf, "p", n_snapshots_total=1024)
    second = exp.interpolated_fields.centers
    exp.synchronize()
    ref = ExportData(g, write_times=[str(i) for i in range(1024)], write_files=False, stream_host=False)
    ref.export(x, a, "p", n_snapshots_total=1024)
    want_a = ref.interpolated_fields.centers.cpu()
    ref.export(x, b, "p", n_snapshots_total=1024)
    want_b = ref.interpolated_fields.centers.cpu()
    assert first.data_ptr() != second.data_ptr()
    assert pt.equal(first, want_a) and pt.equal(second, want_b)
