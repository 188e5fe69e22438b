"""Multi-GPU plumbing: frames shard contiguously over ranks; the only collective is one
all-gather of a uint64 per rank (each rank's compressed byte total), from which every rank
derives where its shard sits in the concatenated stream."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def init_process_group_quiet(device: torch.device) -> None:
    """dist.init_process_group("nccl") plus one warm-up collective (communicator creation is lazy and
    takes seconds), with the process's stdout pointed at stderr meanwhile: NCCL prints its version
    banner with printf, and a benchmark's stdout must stay one JSON line."""
    import os
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group("nccl", device_id=device)
        t = torch.zeros(1, dtype=torch.int64, device=device)
        dist.all_reduce(t)
        torch.cuda.synchronize(device)
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard [lo, hi) of rank `rank`: frames [r*F/W, (r+1)*F/W)."""
    return (rank * n_frames) // world, ((rank + 1) * n_frames) // world


def allgather_totals(total: torch.Tensor) -> torch.Tensor:
    """total: int64[1] on this rank's device (bytes this rank produced) -> int64[world].
    NCCL all-gather of 8 bytes per rank when the process group is NCCL; gloo on CPU tensors."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    out = torch.empty(world, dtype=torch.int64, device=total.device)
    if world == 1:
        out.copy_(total.reshape(1))
        return out
    dist.all_gather_into_tensor(out, total.reshape(1).contiguous())
    return out


def base_offset(all_totals: torch.Tensor, rank: int) -> int:
    return int(all_totals[:rank].sum().item()) if rank else 0


def place_offsets(offsets: torch.Tensor, all_totals: torch.Tensor, rank: int) -> torch.Tensor:
    """Rebase a rank's frame offsets (device int64[n+1]) into the global concatenated stream."""
    if offsets.is_cuda:
        rc = _lib.lib().rspt_gpu_rebase_offsets(offsets.data_ptr(), offsets.numel(), all_totals.data_ptr(), rank,
                                                torch.cuda.current_stream(offsets.device).cuda_stream)
        _lib.check(rc, None, "rspt_gpu_rebase_offsets")
        return offsets
    return offsets + (all_totals[:rank].sum() if rank else 0)
