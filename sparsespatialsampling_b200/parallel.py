"""
Multi-GPU plumbing of the export stage (one process per GPU, ``torch.distributed``).

The interpolation shards naturally: outputs for different snapshots are independent, so every rank owns a contiguous
window of the time axis and needs only the sampled grid and the KNN tables, which rank 0 computes and broadcasts once
(NCCL over NVLink on GPUs, gloo in the CPU tests). There is no collective inside the interpolation itself. Grid
generation is sequential across iterations and is not sharded (replicas only / rank 0 + broadcast).

The SVD of the exported snapshots (utils.py:302-346) is the one step with a real exchange: the Gram matrix contracts
over the cells, so the time-sharded result of the export is first turned into row (cell) shards
(``time_to_row_shards``, one point-to-point block per rank pair), every rank contracts its own rows and the ``T x T``
partial Gram matrices are summed with one all-reduce (``allreduce_sum``), see ``svd.compute_svd_sharded``.
"""
import logging
import os
from typing import List, Optional, Tuple

import torch as pt
import torch.distributed as dist

logger = logging.getLogger(__name__)


def bind_to_gpu_numa_node(device_index: int) -> Optional[List[int]]:
    """
    Pin the calling process to the host cores NVML reports as local to GPU ``device_index`` (its NUMA node).

    The streamed export moves every snapshot batch through pinned host memory at PCIe rate; pinned pages are placed on
    the NUMA node of the allocating thread, so a rank that runs on the other socket pays an inter-socket hop per byte
    in both directions (and, with several ranks, all of them queue on one memory controller). Call this once per
    process BEFORE allocating host buffers -- what ``numactl --cpunodebind`` does for a hand-launched job.
    Returns the core list, or ``None`` when NVML or the affinity call is unavailable (nothing is changed then).
    """
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            index = device_index
            if visible:
                entry = visible.split(",")[device_index].strip()
                if entry.isdigit():
                    index = int(entry)
                    handle = pynvml.nvmlDeviceGetHandleByIndex(index)
                else:
                    handle = pynvml.nvmlDeviceGetHandleByUUID(entry)
            else:
                handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            n_cpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
        local = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = local & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return sorted(allowed)
    except Exception as e:                              # no NVML / not permitted: keep the scheduler's placement
        logger.debug("bind_to_gpu_numa_node: %s", e)
        return None


def snapshot_window(n_snapshots: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous window ``[t0, t1)`` of the time axis owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n_snapshots, world_size)
    t0 = rank * base + min(rank, rem)
    return t0, t0 + base + (1 if rank < rem else 0)


def broadcast_tensors(tensors: List[pt.Tensor], src: int = 0) -> List[pt.Tensor]:
    """In-place broadcast of already allocated, equally shaped tensors from ``src`` to all ranks."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.broadcast(t, src=src)
    return tensors


def broadcast_grid(centers: pt.Tensor, n_dimensions: int, device, src: int = 0) -> pt.Tensor:
    """Share the sampled cell centres ``[Nc, d]`` (fp64) of rank ``src``; other ranks pass ``None``."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return centers
    n = pt.tensor([centers.size(0) if centers is not None else 0], dtype=pt.int64, device=device)
    dist.broadcast(n, src=src)
    if dist.get_rank() != src:
        centers = pt.empty((int(n.item()), n_dimensions), dtype=pt.float64, device=device)
    else:
        centers = centers.to(device=device, dtype=pt.float64).contiguous()
    dist.broadcast(centers, src=src)
    return centers


def _active(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def row_window(n_rows: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous window ``[r0, r1)`` of the cell axis owned by ``rank`` (same rule as ``snapshot_window``)."""
    return snapshot_window(n_rows, world_size, rank)


def time_to_row_shards(data_local: pt.Tensor, n_snapshots_total: int, group=None) -> pt.Tensor:
    """
    Transpose the sharding of an exported field from time windows to cell windows.

    :param data_local: ``[N_cells, T_s]`` or ``[N_cells, D, T_s]``, the snapshot window of this rank
        (``snapshot_window(n_snapshots_total, world, rank)``), all cells
    :param n_snapshots_total: T, the sum of the window lengths over the ranks
    :return: ``[n_local, T]`` / ``[n_local, D, T]``, all snapshots of the cells ``row_window(N_cells, world, rank)``

    Every ordered rank pair exchanges exactly one contiguous block (send rows of the peer x own window, receive own rows
    x window of the peer); the blocks are posted together as one batch of point-to-point operations, which NCCL runs
    as a single grouped all-to-all over NVLink and gloo supports on the CPU.
    """
    if not _active(group):
        if data_local.size(-1) != n_snapshots_total:
            raise ValueError(f"single process holds {data_local.size(-1)} of {n_snapshots_total} snapshots")
        return data_local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_cells = data_local.size(0)
    t0, t1 = snapshot_window(n_snapshots_total, world, rank)
    if data_local.size(-1) != t1 - t0:
        raise ValueError(f"rank {rank} must hold snapshots [{t0}, {t1}), got {data_local.size(-1)} columns")
    r0, r1 = row_window(n_cells, world, rank)
    mid = tuple(data_local.shape[1:-1])
    out = pt.empty((r1 - r0,) + mid + (n_snapshots_total,), dtype=data_local.dtype, device=data_local.device)
    out[..., t0:t1] = data_local[r0:r1]
    ops, landing = [], []
    for q in range(world):
        if q == rank:
            continue
        q0, q1 = row_window(n_cells, world, q)
        s0, s1 = snapshot_window(n_snapshots_total, world, q)
        peer = dist.get_global_rank(group, q) if group is not None else q
        ops.append(dist.P2POp(dist.isend, data_local[q0:q1].contiguous(), peer, group))
        buf = pt.empty((r1 - r0,) + mid + (s1 - s0,), dtype=data_local.dtype, device=data_local.device)
        ops.append(dist.P2POp(dist.irecv, buf, peer, group))
        landing.append((s0, s1, buf))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for s0, s1, buf in landing:
        out[..., s0:s1] = buf
    return out


def allreduce_sum(t: pt.Tensor, group=None) -> pt.Tensor:
    """In-place sum over the ranks (no-op in a single process)."""
    if _active(group):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def gather_rows(local: pt.Tensor, n_rows_total: int, group=None) -> pt.Tensor:
    """Concatenate the row shards ``row_window(n_rows_total, world, rank)`` of all ranks on every rank."""
    if not _active(group):
        return local
    world = dist.get_world_size(group)
    widths = [row_window(n_rows_total, world, q) for q in range(world)]
    n_max = max(q1 - q0 for q0, q1 in widths)            # windows differ by at most one row: pad to equal blocks
    mine = pt.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    mine[:local.size(0)] = local
    parts = [pt.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return pt.cat([p[:q1 - q0] for p, (q0, q1) in zip(parts, widths)], dim=0)
