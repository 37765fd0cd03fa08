"""
world_size-2 gloo test of the N>1 path (host logic only, no GPU): rank 0 owns the KNN tables and broadcasts them,
every rank interpolates its own snapshot window, the concatenation equals the single-process result.
"""
import os
import socket

import numpy as np
import torch as pt
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import s3_oracle as orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparsespatialsampling_b200.parallel import snapshot_window, broadcast_tensors, broadcast_grid
    rng = np.random.default_rng(0)                     # same data on every rank
    N, Nc, k, T = 400, 150, 8, 11
    data = rng.standard_normal((N, 2, T)).astype(np.float32)
    if rank == 0:
        idx = pt.from_numpy(rng.integers(0, N, (Nc, k)))
        w = pt.from_numpy(rng.random((Nc, k)))
        centers = pt.from_numpy(rng.random((Nc, 2)))
    else:
        idx, w, centers = pt.zeros((Nc, k), dtype=pt.int64), pt.zeros((Nc, k), dtype=pt.float64), None
    centers = broadcast_grid(centers, 2, "cpu", src=0)
    broadcast_tensors([idx, w], src=0)
    t0, t1 = snapshot_window(T, world, rank)
    out = orc.interpolate(w.numpy(), idx.numpy(), data[:, :, t0:t1])
    np.save(os.path.join(tmp, f"part{rank}.npy"), out)
    np.save(os.path.join(tmp, f"centers{rank}.npy"), centers.numpy())
    if rank == 0:
        np.save(os.path.join(tmp, "full.npy"), orc.interpolate(w.numpy(), idx.numpy(), data))
    dist.barrier()
    dist.destroy_process_group()


def test_snapshot_sharded_interpolation_world_size_2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f"part{r}.npy") for r in range(world)]
    full = np.load(tmp_path / "full.npy")
    assert np.array_equal(np.concatenate(parts, axis=2), full)
    assert np.array_equal(np.load(tmp_path / "centers0.npy"), np.load(tmp_path / "centers1.npy"))


def _svd_exchange_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sparsespatialsampling_b200.parallel import (snapshot_window, row_window, time_to_row_shards, allreduce_sum,
                                                     gather_rows)
    rng = np.random.default_rng(1)                     # same data on every rank
    Nc, D, T = 37, 2, 13                               # odd sizes: unequal windows on both axes
    full = rng.standard_normal((Nc, D, T)).astype(np.float32)
    vol = rng.random(Nc) + 0.5
    t0, t1 = snapshot_window(T, world, rank)
    rows = time_to_row_shards(pt.from_numpy(full[:, :, t0:t1].copy()), T)
    r0, r1 = row_window(Nc, world, rank)
    assert np.array_equal(rows.numpy(), full[r0:r1])
    rows2d = time_to_row_shards(pt.from_numpy(full[:, 0, t0:t1].copy()), T)
    assert np.array_equal(rows2d.numpy(), full[r0:r1, 0])
    # local weighted Gram of the own rows (oracle arithmetic), summed over the ranks == Gram of the whole matrix
    g = pt.from_numpy(orc.weighted_gram(rows.numpy().reshape(-1, T), np.repeat(vol[r0:r1], D)))
    allreduce_sum(g)
    everything = gather_rows(rows, Nc)
    assert np.array_equal(everything.numpy(), full)
    np.save(os.path.join(tmp, f"gram{rank}.npy"), g.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_svd_exchange_world_size_3(tmp_path):
    """time -> cell re-sharding, all-reduce of the partial Gram matrices and the row gather of compute_svd_sharded."""
    world = 3
    mp.spawn(_svd_exchange_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(1)
    full = rng.standard_normal((37, 2, 13)).astype(np.float32)
    vol = rng.random(37) + 0.5
    ref = orc.weighted_gram(full.reshape(-1, 13), np.repeat(vol, 2))
    for r in range(world):
        g = np.load(tmp_path / f"gram{r}.npy")
        assert np.allclose(g, ref, rtol=1e-12, atol=1e-12)
        assert np.array_equal(g, np.load(tmp_path / "gram0.npy"))


def test_bind_to_gpu_numa_node_is_harmless_without_nvml():
    # no GPU / NVML in the CPU container: the helper must leave the affinity alone and say so
    from sparsespatialsampling_b200.parallel import bind_to_gpu_numa_node
    before = os.sched_getaffinity(0)
    cores = bind_to_gpu_numa_node(0)
    assert cores is None or set(cores) <= before
    if cores is None:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)


def test_row_and_snapshot_windows_partition_the_axis():
    from sparsespatialsampling_b200.parallel import snapshot_window, row_window
    for n, world in [(1000, 8), (13, 3), (5, 8), (0, 2)]:
        edges = [snapshot_window(n, world, r) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
        sizes = [b - a for a, b in edges]
        assert max(sizes) - min(sizes) <= 1
        assert edges == [row_window(n, world, r) for r in range(world)]
