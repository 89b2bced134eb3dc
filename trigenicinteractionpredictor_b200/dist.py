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


def broadcast_from_first_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In place: every rank of `group` receives the tensor of the group's first rank; a no-op in a single process."""
    if world_size(group) > 1:
        src = td.get_global_rank(group, 0) if group is not None and group is not td.group.WORLD else 0
        td.broadcast(t, src=src, group=group)
    return t


def broadcast_int(value: int, group=None) -> int:
    """The integer rank 0 holds, on every rank (e.g. the seed of a link-sharded run)."""
    if world_size(group) == 1:
        return int(value)
    box = [int(value)]
    td.broadcast_object_list(box, src=td.get_global_rank(group, 0) if group is not None and group is not td.group.WORLD else 0,
                             group=group)
    return int(box[0])


def barrier(group=None):
    if world_size(group) > 1:
        td.barrier(group=group)


def max_over_ranks(x: float, device=None, group=None) -> float:
    if world_size(group) == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    td.all_reduce(t, op=td.ReduceOp.MAX, group=group)
    return float(t.item())


class PeerExchange:
    """Statistics exchange over NVLink peer memory instead of an NCCL allreduce.  Each rank owns one device buffer
    `[2][n_pad] statistics | theta [n_theta] | p [n_p] | 2 * world uint64 flags`; CUDA IPC handles are swapped once
    through the process group, after which every rank holds a device pointer to every peer's buffer.
      mode "push" (default): tip_peer_push_mstep - every rank stores its statistics into its slot of every peer's inbox, one
                           handshake, local sum in rank order + M-step, one kernel; inboxes double-buffered (i & 1);
      mode "rs":           tip_peer_mstep - reduce-scatter of the statistics, M-step of the own slice, all-gather of the
                           new theta / p by remote stores, one kernel, two handshakes; uses statistics buffer 0 only (theta
                           and p of the engine live INSIDE the shared buffer so that peers can write them);
      mode "gather":       tip_peer_barrier + tip_normalise_peers - every rank reads all n buffers (round 1); iteration i
                           uses statistics buffer i & 1 (double buffering makes one barrier per iteration sufficient)."""

    def __init__(self, n_stats: int, device, group, n_theta: int = 0, n_p: int = 0, mode: str = "push"):
        import ctypes
        from . import _cabi
        self.lib = _cabi.load()
        self.group = group
        self.mode = mode
        self.world, self.rank = td.get_world_size(group), td.get_rank(group)
        self.n_stats = n_stats
        self.n_pad = (n_stats + 31) // 32 * 32
        self.n_theta, self.n_p = n_theta, n_p
        th_pad, p_pad = (n_theta + 31) // 32 * 32, (n_p + 31) // 32 * 32
        self.off_theta = 2 * self.n_pad
        self.off_p = self.off_theta + th_pad
        self.off_flags = self.off_p + p_pad
        self.off_inbox = (self.off_flags + 2 * self.world + 31) // 32 * 32
        n_inbox = 2 * self.world * self.n_pad if mode == "push" else 0
        self.buf = torch.zeros(self.off_inbox + n_inbox, dtype=torch.float64, device=device)
        self.epoch = torch.zeros(4, dtype=torch.int64, device=device)      # {epoch, CTAs done, timed out} of tip_peer_mstep
        handle = (ctypes.c_char * 64)()
        off = ctypes.c_int64(0)
        _cabi.check(self.lib.tip_ipc_export(ctypes.c_void_p(self.buf.data_ptr()), handle, ctypes.byref(off)),
                    "tip_ipc_export")
        mine = (bytes(handle.raw), int(off.value))
        everyone = [None] * self.world
        td.all_gather_object(everyone, mine, group=group)
        bases = []
        for r, (h, o) in enumerate(everyone):
            if r == self.rank:
                bases.append(self.buf.data_ptr())
            else:
                out = ctypes.c_void_p(0)
                hb = ctypes.create_string_buffer(h, 64)
                _cabi.check(self.lib.tip_ipc_import(hb, ctypes.c_int64(o), ctypes.byref(out)), "tip_ipc_import")
                bases.append(out.value)
        arr = ctypes.c_void_p * self.world
        self.stats_ptrs = [arr(*[b + i * self.n_pad * 8 for b in bases]) for i in range(2)]
        self.theta_ptrs = arr(*[b + self.off_theta * 8 for b in bases])
        self.p_ptrs = arr(*[b + self.off_p * 8 for b in bases])
        self.flag_ptrs = arr(*[b + self.off_flags * 8 for b in bases])
        self.inbox_ptrs = [arr(*[b + (self.off_inbox + i * self.world * self.n_pad) * 8 for b in bases]) for i in range(2)]
        torch.cuda.synchronize(device)
        td.barrier(group=group)          # every rank has mapped every buffer before anyone signals

    def stats(self, parity: int) -> torch.Tensor:
        return self.buf[parity * self.n_pad: parity * self.n_pad + self.n_stats]

    def theta(self) -> torch.Tensor:
        return self.buf[self.off_theta: self.off_theta + self.n_theta]

    def p(self) -> torch.Tensor:
        return self.buf[self.off_p: self.off_p + self.n_p]

    def check(self):
        """Raises if a barrier timed out (a peer stopped participating)."""
        if int(self.epoch[0].item()) == -1:
            raise RuntimeError("the peer exchange timed out: a peer rank stopped participating; the statistics of that "
                               "iteration were incomplete and the parameters since then are invalid")
