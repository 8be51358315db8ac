"""Multi-GPU layout of the path: videos are independent (frames fold into the batch,
src/models/video_encoder.py:293-294; the temporal mean is per video, :256-258; decoder rows are
independent), so the batch shards by video with replicated weights and NO data-path collective.
The only exchange is gathering the generated token ids — one all_gather of
int32 [B_local, max_new] + lengths per batch, latency-bound over NVLink.

One process per GPU (torchrun); works with the nccl backend on GPUs and with gloo on CPU
tensors (used by the world_size-2 CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_videos: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of global video indices owned by `rank`: v -> rank v // ceil(n/world).
    Ragged tails are allowed (the last ranks may own fewer, possibly zero, videos)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = (n_videos + world - 1) // world
    lo = min(rank * per, n_videos)
    hi = min(lo + per, n_videos)
    return lo, hi


def gather_ids(ids: torch.Tensor, lengths: torch.Tensor, n_videos: int, group=None):
    """all_gather the per-rank token ids/lengths and return them in GLOBAL video order.
    ids int32 [B_local, max_new] (eos padded), lengths int32 [B_local]; every rank gets the result."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return ids, lengths
    per = (n_videos + world - 1) // world
    max_new = ids.shape[1]
    pad_ids = torch.zeros(per, max_new, dtype=ids.dtype, device=ids.device)
    pad_len = torch.zeros(per, dtype=lengths.dtype, device=lengths.device)
    pad_ids[: ids.shape[0]] = ids
    pad_len[: lengths.shape[0]] = lengths
    all_ids = [torch.empty_like(pad_ids) for _ in range(world)]
    all_len = [torch.empty_like(pad_len) for _ in range(world)]
    dist.all_gather(all_ids, pad_ids, group=group)
    dist.all_gather(all_len, pad_len, group=group)
    out_ids, out_len = [], []
    for r in range(world):
        lo, hi = shard_range(n_videos, world, r)
        out_ids.append(all_ids[r][: hi - lo])
        out_len.append(all_len[r][: hi - lo])
    return torch.cat(out_ids, 0), torch.cat(out_len, 0)
