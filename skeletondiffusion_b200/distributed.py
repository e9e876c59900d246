"""Multi-GPU plumbing: one process per GPU, windows sharded across ranks, no collective inside the
sampling loop; a single all-gather of per-window metric vectors at the end."""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist

from .pipeline import shard_windows

__all__ = ["init_from_env", "gather_window_metrics"]


def init_from_env() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from torchrun's environment; initialises NCCL when world_size > 1."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    return rank, local_rank, world


def gather_window_metrics(local: Dict[str, torch.Tensor], num_windows: int, rank: int, world_size: int) -> Dict[str, torch.Tensor]:
    """All-gather per-window metric vectors (e.g. ADE/FDE/APD, 4 bytes per window each) into global window order.
    Shards may differ by one window, so each rank pads to the largest shard."""
    if world_size == 1:
        return dict(local)
    spans = [shard_windows(num_windows, r, world_size) for r in range(world_size)]
    width = max(hi - lo for lo, hi in spans)
    keys = sorted(local)
    ref = local[keys[0]]
    packed = torch.zeros(len(keys), width, dtype=torch.float32, device=ref.device)
    lo, hi = spans[rank]
    for i, k in enumerate(keys):
        packed[i, :hi - lo] = local[k].to(torch.float32)
    bufs = [torch.empty_like(packed) for _ in range(world_size)]
    dist.all_gather(bufs, packed)
    out = {}
    for i, k in enumerate(keys):
        out[k] = torch.cat([bufs[r][i, :spans[r][1] - spans[r][0]] for r in range(world_size)])
    return out
