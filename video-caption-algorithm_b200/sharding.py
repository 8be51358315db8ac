"""Multi-GPU layout of the path: videos are independent (frames fold into the batch,
src/models/video_encoder.py:293-294; the temporal mean is per video, :256-258; decoder rows are
independent), so the batch shards by video with replicated weights and NO data-path collective.
The only exchange is gathering the generated token ids — ONE all_gather per batch of a
packed int32 [B_local, max_new + 1] block (ids | length) into a preallocated [world, B_local, max_new + 1]
buffer (`IdGatherer`), latency-bound over NVLink.

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


class IdGatherer:
    """The per-batch id exchange with nothing allocated per call: a packed send block [per, max_new + 1] (ids, then the length in
    the last column) and a receive buffer [world, per, max_new + 1], one `all_gather_into_tensor` per batch on the caller's stream.
    `per` = videos per rank (equal on every rank; ragged global batches pad the last ranks' blocks with length 0)."""

    def __init__(self, per_rank: int, max_new: int, world: int, device, group=None):
        self.per, self.max_new, self.world, self.group = int(per_rank), int(max_new), int(world), group
        self.send = torch.zeros(self.per, self.max_new + 1, dtype=torch.int32, device=device)
        self.recv = torch.zeros(self.world, self.per, self.max_new + 1, dtype=torch.int32, device=device)

    def gather(self, ids: torch.Tensor, lengths: torch.Tensor) -> torch.Tensor:
        """ids int32 [b, max_new], lengths int32 [b], b <= per.  Returns the receive buffer [world, per, max_new + 1]
        (block r = rank r's rows; valid until the next call)."""
        b = ids.shape[0]
        self.send[:b, : self.max_new].copy_(ids)
        self.send[:b, self.max_new].copy_(lengths)
        if b < self.per:
            self.send[b:].zero_()
        if self.world == 1:
            self.recv[0].copy_(self.send)
        else:
            dist.all_gather_into_tensor(self.recv.view(-1), self.send.view(-1), group=self.group)
        return self.recv

    def check_own_block(self, gathered: torch.Tensor, ids: torch.Tensor, lengths: torch.Tensor, rank: int) -> bool:
        """block `rank` of a gathered buffer holds exactly this rank's ids and lengths (one host sync: call outside timed regions)."""
        b = ids.shape[0]
        blk = gathered[rank]
        return bool(torch.equal(blk[:b, : self.max_new], ids.to(torch.int32)) and torch.equal(blk[:b, self.max_new], lengths.to(torch.int32)))

    def global_order(self, gathered: torch.Tensor, n_videos: int):
        """(ids [n_videos, max_new], lengths [n_videos]) in global video order from a gathered buffer."""
        flat = gathered.reshape(self.world * self.per, self.max_new + 1)
        rows = []
        for r in range(self.world):
            lo, hi = shard_range(n_videos, self.world, r)
            rows.append(flat[r * self.per: r * self.per + (hi - lo)])
        out = torch.cat(rows, 0)
        return out[:, : self.max_new], out[:, self.max_new]
