"""
compute_svd_sharded on 2 GPUs (one process per GPU, NCCL): time-sharded export result -> cell shards -> local tcgen05
Gram -> all-reduce -> identical factors on every rank. Compared with the single-GPU compute_svd and with the oracle
(torch.linalg.svd of the weighted matrix, oracle/s3_oracle.py:compute_svd). Skipped on a single-GPU box.
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


def _field(nc, d, t):
    rng = np.random.default_rng(5)
    x = rng.random(nc)
    tt = np.linspace(0, 1, t)
    modes = [np.sin(2 * np.pi * (k + 1) * x)[:, None] * np.cos(2 * np.pi * (k + 1) * tt + k)[None] / (k + 1) for k in range(6)]
    base = sum(modes) + 0.01 * rng.standard_normal((nc, t))
    if d == 0:
        return base.astype(np.float32), rng.random(nc) + 0.5
    return np.stack([base * (j + 1) + 0.1 * j for j in range(d)], axis=1).astype(np.float32), rng.random(nc) + 0.5


def _worker(rank, world, port, tmp, d):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    pt.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=pt.device("cuda", rank))
    from sparsespatialsampling_b200.parallel import snapshot_window, row_window
    from sparsespatialsampling_b200.svd import compute_svd, compute_svd_sharded
    nc, t, r = 5003, 301, 8
    full, vol = _field(nc, d, t)
    t0, t1 = snapshot_window(t, world, rank)
    local = pt.from_numpy(np.ascontiguousarray(full[..., t0:t1])).cuda()
    s, u, v = compute_svd_sharded(local, pt.from_numpy(vol), rank=r, sharded_by="time", n_snapshots_total=t,
                                  gather_modes=True)
    r0, r1 = row_window(nc, world, rank)
    s2, u2, v2 = compute_svd_sharded(pt.from_numpy(full[r0:r1]).cuda(), pt.from_numpy(vol[r0:r1]), rank=r)
    assert pt.equal(s, s2) and pt.equal(v, v2) and pt.equal(u[r0:r1], u2)
    if rank == 0:
        s1, u1, v1 = compute_svd(pt.from_numpy(full).cuda(), pt.from_numpy(vol), rank=r)
        np.savez(os.path.join(tmp, "out.npz"), s=s.cpu().numpy(), u=u.cpu().numpy(), v=v.cpu().numpy(),
                 s1=s1.cpu().numpy(), u1=u1.cpu().numpy(), v1=v1.cpu().numpy())
    else:
        np.savez(os.path.join(tmp, "out1.npz"), s=s.cpu().numpy(), v=v.cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("d", [0, 2])
def test_sharded_svd_two_gpus(tmp_path, d):
    if pt.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import s3_oracle as orc
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path), d), nprocs=2, join=True)
    o, o1 = np.load(tmp_path / "out.npz"), np.load(tmp_path / "out1.npz")
    assert np.array_equal(o["s"], o1["s"]) and np.array_equal(o["v"], o1["v"])     # identical on every rank
    full, vol = _field(5003, d, 301)
    s_ref, u_ref, v_ref = orc.compute_svd(full, vol, 8)
    tol = 1e-4 * s_ref[0]
    assert np.abs(o["s"] - s_ref).max() <= tol
    assert np.abs(o["s"] - o["s1"]).max() <= tol
    for i in range(6):                                 # well separated modes, up to sign
        a, b = o["u"][..., i].reshape(-1), np.asarray(u_ref)[..., i].reshape(-1)
        assert abs(a @ b) / (np.linalg.norm(a) * np.linalg.norm(b)) >= 1 - 1e-4
        assert abs(o["v"][:, i] @ np.asarray(v_ref)[:, i]) >= 1 - 1e-4
        a1 = o["u1"][..., i].reshape(-1)
        assert abs(a @ a1) / (np.linalg.norm(a) * np.linalg.norm(a1)) >= 1 - 1e-6
