"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests).

Two partitionings of the reference's work:
  * link shards   - theta / p replicated, contiguous slices of the link list per rank, ONE allreduce
                    (sum, fp64) of the statistics buffer per EM iteration between E-step and M-step;
  * sample shards - independent random restarts (what run.sh does with GNU parallel): no communication.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as td


def is_initialized() -> bool:
    return td.is_available() and td.is_initialized()


def world_size(group=None) -> int:
    return td.get_world_size(group) if is_initialized() else 1


def rank(group=None) -> int:
    return td.get_rank(group) if is_initialized() else 0


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Initialise from torchrun's environment.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            td.init_process_group(backend, rank=rk, world_size=world, device_id=torch.device("cuda", local))
        else:
            td.init_process_group(backend, rank=rk, world_size=world)
    return rk, world, local


def shard_bounds(n: int, rk: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of n items for rank rk; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rk * base + min(rk, rem)
    return lo, lo + base + (1 if rk < rem else 0)


def samples_for_rank(sample_ini: int, num_samples: int, rk: int, world: int) -> list[int]:
    """Round-robin assignment of restart samples to ranks (7/7/6/... for 50 samples on 8 GPUs)."""
    return [s for s in range(sample_ini, sample_ini + num_samples) if (s - sample_ini) % world == rk]


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum over ranks; a no-op in a single process."""
    if world_size(group) > 1:
        td.all_reduce(t, op=td.ReduceOp.SUM, group=group)
    return t


def barrier(group=None):
    if world_size(group) > 1:
        td.barrier(group=group)


def max_over_ranks(x: float, device=None, group=None) -> float:
    if world_size(group) == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    td.all_reduce(t, op=td.ReduceOp.MAX, group=group)
    return float(t.item())
