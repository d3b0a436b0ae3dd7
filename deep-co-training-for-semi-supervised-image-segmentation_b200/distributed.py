"""Data-parallel glue for the consistency path (one process per GPU, torch.distributed).

The per-pixel path needs no exchange: the batch is sharded on dim 0 and
the only cross-rank coupling is a handful of scalars / small integer tensors:

  * the global mean of a loss map:  all_reduce(SUM) of [sum_local, n_local]
  * Dice '3d' counts (summed over the whole batch) and confusion matrices: all_reduce(SUM) of int64

These helpers are backend-agnostic (NCCL on the B200 box, gloo in the CPU tests) and are the
ONLY collectives this package issues; gradient all-reduce belongs to DDP around the networks.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def shard_batch(n_items: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    """[start, stop) of this rank's contiguous slice of a batch of ``n_items`` (equal shards, remainder to
    the first ranks) -- the same split for the unlabeled batch and every labeled batch."""
    if rank is None:
        rank = dist.get_rank() if is_distributed() else 0
    if world is None:
        world = dist.get_world_size() if is_distributed() else 1
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def global_mean(local_sum: torch.Tensor, n_local: int, group=None) -> torch.Tensor:
    """sum_local / n_local over all ranks == the reference's ``.mean()`` over the un-sharded batch."""
    buf = torch.stack([local_sum.reshape(()).to(torch.float64),
                       torch.tensor(float(n_local), dtype=torch.float64, device=local_sum.device)])
    if is_distributed():
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return (buf[0] / buf[1]).to(torch.float32)


def global_pixel_count(n_local: int, device, group=None) -> int:
    """Total pixel count over ranks (the ``n_global`` of the fused losses)."""
    if not is_distributed():
        return int(n_local)
    t = torch.tensor([n_local], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def all_reduce_counts(counts: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM of integer count tensors (Dice '3d' [.,C,3] after a batch sum, confusion [C,C])."""
    if is_distributed():
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def all_gather_rows(rows: torch.Tensor, group=None) -> torch.Tensor:
    """Concatenate per-image Dice rows ('2d') from all ranks in rank order (equal row counts)."""
    if not is_distributed():
        return rows
    out = [torch.empty_like(rows) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, rows.contiguous(), group=group)
    return torch.cat(out, 0)
