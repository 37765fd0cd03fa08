"""
Sharded export on 2 GPUs (one process per GPU, NCCL): ExportData(distributed=True) -- rank 0 builds the KNN tables and
broadcasts them, every rank interpolates and writes its own snapshot window (host batches through the streamed path and
device batches), rank 0 writes the XDMF. The concatenated shards are bit-identical to the single-GPU export and within
1e-5 of the CPU oracle (export.py:279-313, examples/s3_for_cylinder3D_Re3900.py:28-69). Skipped on a single-GPU box.
"""
import os
import socket

import numpy as np
import pytest
import torch as pt
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Grid:
    pass


def _case(tmp):
    rng = np.random.default_rng(3)
    n, nc, t = 30000, 9001, 203
    x = pt.from_numpy(rng.random((n, 2)))
    g = _Grid()
    g.n_dimensions, g.size_initial_cell = 2, 1.0
    g.centers = pt.from_numpy(rng.random((nc, 2)))
    g.vertices = pt.from_numpy(rng.random((nc + 50, 2)))
    g.faces = pt.from_numpy(rng.integers(0, nc + 50, (nc, 4)).astype(np.int32))
    g.levels = pt.from_numpy(rng.integers(1, 5, (nc, 1)))
    g.metric = pt.from_numpy(rng.random(n))
    g.save_path, g.save_name, g.grid_name = tmp, "run", "grid"
    u = pt.from_numpy(rng.standard_normal((n, 2, t)).astype(np.float32))
    p = pt.from_numpy(rng.standard_normal((n, 1, t)).astype(np.float32))
    return x, g, u, p, [f"{0.01 * i:.2f}" for i in range(t)]


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    pt.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=pt.device("cuda", rank))
    from sparsespatialsampling_b200.export import ExportData
    from sparsespatialsampling_b200.parallel import snapshot_window
    x, g, u, p, times = _case(tmp)
    t = len(times)
    exp = ExportData(g, write_times=times, distributed=True)
    exp._stream_min_elements = 0
    t0, t1 = snapshot_window(t, world, rank)
    mid = (t0 + t1) // 2
    exp.export(x, u[:, :, t0:mid].contiguous().pin_memory(), "U", n_snapshots_total=t)     # host batch: streamed path
    exp.export(x, u[:, :, mid:t1].cuda(), "U", n_snapshots_total=t)                        # device batch
    exp.export(x, p[:, :, t0:t1].contiguous(), "p", n_snapshots_total=t)                   # pageable host batch
    assert exp._tables_centers.n == g.centers.size(0)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_export_two_gpus(tmp_path):
    if pt.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import s3_oracle as orc
    from sparsespatialsampling_b200.data import Dataloader
    from sparsespatialsampling_b200.export import ExportData
    tmp = str(tmp_path)
    mp.spawn(_worker, args=(2, _free_port(), tmp), nprocs=2, join=True)
    assert {"run.h5", "run.part1.h5", "run.xdmf"} <= set(os.listdir(tmp))
    x, g, u, p, times = _case(os.path.join(tmp, "single"))
    one = ExportData(g, write_times=times)
    one.export(x, u.cuda(), "U")
    idx = one._knn_idx_centers.cpu().numpy().astype(np.int64)
    w = one._knn_w_centers.cpu().numpy()
    one.export(x, p.cuda(), "p")
    many, single = Dataloader(tmp, "run.h5"), Dataloader(os.path.join(tmp, "single"), "run.h5")
    assert many.write_times == single.write_times and len(many.write_times) == len(times)
    for name, field in (("U", u), ("p", p)):
        got, want = many.load_snapshot(name), single.load_snapshot(name)
        assert pt.equal(got, want)                                                         # shards == single GPU, bit for bit
        ref = orc.interpolate(w, idx, field.numpy())
        scale = np.abs(field.numpy()[idx]).max(axis=1)
        got = got.numpy().reshape(ref.shape)
        assert (np.abs(got - ref) <= 1e-5 * np.maximum(scale, 1e-30)).all()
    assert pt.equal(many.metric, single.metric)
    xdmf = open(os.path.join(tmp, "run.xdmf")).read()
    assert xdmf.count("<Time Value=") == len(times) and "run.part1.h5:/data/2.02/U_center" in xdmf
