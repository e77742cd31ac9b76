"""Multi-GPU plumbing: contiguous shards of the batch, one process per GPU, results gathered on demand.

The hot path has no exchange step: every configuration and every edge is independent, the chain /
scene descriptors (< 8 KB) are replicated on each GPU, and rank r evaluates the contiguous slice
`[r*N/G, (r+1)*N/G)`.  NCCL (or gloo on CPU, for the tests) is used only to
gather verdict bytes or poses when a caller wants them in one place; for BASELINE config 4 that is
1 MB for 1 M edges.
"""
import os

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, exhaustive, non-overlapping: shard sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return (rank * n) // world, ((rank + 1) * n) // world


def init_from_env(backend: str | None = None):
    """Join the process group described by RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun).

    Returns (rank, local_rank, world_size).  Single-process runs (no RANK in the env) skip the group.
    """
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def gather_shards(local: torch.Tensor, n_total: int, dst: int | None = None):
    """Concatenate the per-rank slices of a sharded result along dim 0.

    `local` is this rank's slice (shard_bounds order).  dst=None -> every rank gets the full tensor
    (all_gather); dst=r -> only rank r does, the others get None.  Shards may differ by one row, so
    they are padded to the largest shard for the collective and trimmed afterwards.
    """
    if not dist.is_initialized() or dist.get_world_size() == 1:
        if local.shape[0] != n_total:
            raise ValueError("single-process gather: local slice must be the whole result")
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: slice has {local.shape[0]} rows, expected {sizes[rank]}")
    pad = max(sizes)
    buf = local
    if local.shape[0] != pad:
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    buf = buf.contiguous()
    as_bool = buf.dtype == torch.bool
    if as_bool:
        buf = buf.to(torch.uint8)
    if dst is None:
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
    else:
        out = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, out, dst=dst)
        if rank != dst:
            return None
    full = torch.cat([o[:s] for o, s in zip(out, sizes)], dim=0)
    return full.to(torch.bool) if as_bool else full


def max_over_ranks(value: float, device=None) -> float:
    """Timing rule for multi-GPU numbers: the slowest rank defines the step."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
