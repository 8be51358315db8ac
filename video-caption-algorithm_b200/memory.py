"""Device memory owned by the B200 path: the bf16 KV cache and a workspace arena.

The reference's `core/memory.py` (MemoryManager.snapshot/cleanup/oom_guard, :19-46)
only watches `mem_get_info` and empties the allocator cache on OOM; its KV cache is
HF `DynamicCache`, which re-allocates every layer's K/V with `torch.cat` each step
(SURVEY.md §2.2b).  Here the cache is one contiguous, preallocated bf16 tensor
`[layers, 2, n_seq, heads, s_max, 64]` written in place by the attention kernel,
with an optional int32 slot table `[n_seq, s_max]` so beam reorder is an index
update instead of a copy.  All scratch lives in buffers that only ever grow, so a
steady-state batch performs no allocation (and CUDA graphs keep valid pointers).
"""
from __future__ import annotations

import ctypes as C
from contextlib import contextmanager
from dataclasses import dataclass

import torch

from . import lib as L


@dataclass(frozen=True)
class GpuMemorySnapshot:
    """Same fields as the reference's snapshot (core/memory.py:11-16)."""
    allocated_mb: float
    reserved_mb: float
    free_mb: float | None = None
    total_mb: float | None = None


class MemoryManager:
    """core/memory.py:19-46 surface, kept so engine-level callers need no change."""

    def __init__(self, device):
        self.device = torch.device(device)

    def snapshot(self) -> GpuMemorySnapshot:
        free_b, total_b = torch.cuda.mem_get_info(self.device)
        return GpuMemorySnapshot(torch.cuda.memory_allocated(self.device) / 2**20, torch.cuda.memory_reserved(self.device) / 2**20,
                                 free_b / 2**20, total_b / 2**20)

    def cleanup(self):
        torch.cuda.empty_cache()

    @contextmanager
    def oom_guard(self):
        try:
            yield
        except torch.cuda.OutOfMemoryError:
            self.cleanup()
            raise


class KvCache:
    """Contiguous bf16 KV cache; `length` = positions already written."""

    def __init__(self, layers: int, n_seq: int, heads: int, s_max: int, head_dim: int, device, with_slots: bool = False):
        self.kv = torch.empty(layers, 2, n_seq, heads, s_max, head_dim, device=device, dtype=torch.bfloat16)
        self.slot = None
        if with_slots:
            self.slot = torch.arange(n_seq, device=device, dtype=torch.int32).view(n_seq, 1).repeat(1, s_max).contiguous()
        self.c = L.VcKvCache()
        self.c.kv = self.kv.data_ptr()
        self.c.slot = self.slot.data_ptr() if self.slot is not None else None
        self.c.layers, self.c.n_seq, self.c.heads, self.c.s_max, self.c.head_dim = layers, n_seq, heads, s_max, head_dim
        self.length = 0
        self.s_max = s_max
        self.n_seq = n_seq

    def bytes(self) -> int:
        return self.kv.numel() * 2

    # HF cache API used by the benchmark loop only as an opaque token
    def get_seq_length(self) -> int:
        return self.length


class Workspace:
    """Named grow-only device buffers."""

    def __init__(self, model):
        self._m = model
        self._buf: dict[str, torch.Tensor] = {}

    def _get(self, name: str, numel: int, dtype) -> torch.Tensor:
        t = self._buf.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(max(numel, 1), device=self._m.device, dtype=dtype)
            self._buf[name] = t
        return t

    def patches(self, n_frames: int) -> torch.Tensor:
        d = self._m.dims
        rows = n_frames * (d["tokens"] - 1)
        return self._get("patches", rows * d["k_pad"], torch.bfloat16)

    def cls(self, n_frames: int) -> torch.Tensor:
        return self._get("cls", n_frames * self._m.dims["vit_dim"], torch.float32)

    def vit(self, chunk_frames: int) -> torch.Tensor:
        need = L.load().vc_vit_workspace_bytes(C.byref(self._m.packed.vit), chunk_frames)
        return self._get("vit", need, torch.uint8)

    def gpt(self, n_seq: int, rows: int) -> torch.Tensor:
        need = L.load().vc_gpt_workspace_bytes(C.byref(self._m.packed.gpt), n_seq, rows)
        return self._get("gpt", need, torch.uint8)

    def frames(self, shape) -> torch.Tensor:
        n = 1
        for s in shape:
            n *= int(s)
        return self._get("frames", n, torch.uint8)[:n].view(*shape)

    def named(self, name: str, numel: int, dtype) -> torch.Tensor:
        return self._get(name, numel, dtype)

    def total_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self._buf.values())
