"""
Multi-GPU plumbing of the export stage (one process per GPU, ``torch.distributed``).

The interpolation shards naturally: outputs for different snapshots are independent, so every rank owns a contiguous
window of the time axis and needs only the sampled grid and the KNN tables, which rank 0 computes and broadcasts once
(NCCL over NVLink on GPUs, gloo in the CPU tests). There is no collective inside the interpolation itself. Grid
generation is sequential across iterations and is not sharded (replicas only / rank 0 + broadcast).
"""
from typing import List, Tuple

import torch as pt
import torch.distributed as dist


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
