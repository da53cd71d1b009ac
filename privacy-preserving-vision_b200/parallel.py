"""Data-parallel plumbing for the camera: one process per GPU, batch sharded over ranks.

The optical model shards trivially over images: every rank synthesises the PSF
locally from the same height map and convolves its own slice of the batch; the only exchange is the
all-reduce of dL/dh (N*N floats - 256 KB at N=256) after the PSF chain's backward, which is linear in
dL/dpsf, so reducing after it moves 3x fewer bytes than reducing dL/dpsf.

These helpers are backend agnostic (NCCL on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard", "allreduce_height_grad", "init_from_env", "PeerComm"]


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of a global batch owned by `rank`; sizes differ by at most one image."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_range(t.shape[0], rank, world)
    return t[lo:hi]


def allreduce_height_grad(grad_h: torch.Tensor, group=None, average: bool = True) -> torch.Tensor:
    """In-place all-reduce(SUM) of dL/dh over `group`; with `average` the result is divided by the world
    size (loss = mean over ranks of the per-rank losses: per-rank regulariser terms then count once)."""
    if not dist.is_available() or not dist.is_initialized():
        return grad_h
    dist.all_reduce(grad_h, op=dist.ReduceOp.SUM, group=group)
    if average:
        grad_h /= dist.get_world_size(group)
    return grad_h


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """torchrun-style init (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*); returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


class PeerComm:
    """NVLink peer-memory buffers for the fused dL/dh all-reduce (``b200cam_psf_bwd_allreduce``).

    One symmetric allocation per rank (``torch.distributed._symmetric_memory``: CUDA VMM handles exchanged over the
    process group's store, every rank's buffer mapped into every process); the library's last PSF-backward kernel pushes
    its rows of dL/dh straight into the peers' buffers and sums them - no NCCL call on the data path.
    Construction is a collective.  Raises if symmetric memory is not available (callers fall back to NCCL)."""

    def __init__(self, N: int, device: torch.device, group=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        group = group or dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        lib = _lib.load_library()
        nbytes = lib.b200cam_comm_bytes(N, self.world)
        if nbytes == 0:
            raise RuntimeError(f"b200cam: no peer all-reduce for N={N}, world={self.world}")
        self.buf = symm_mem.empty(nbytes // 4, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier(group)                    # every rank's flags are zero before anyone pushes
        ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self.ptr_array = (ctypes.c_void_p * self.world)(*ptrs)
        self.N = N
