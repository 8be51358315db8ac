"""Host side of the frame resize: Pillow's coefficient rule, evaluated once per (input size, output size).

The reference resizes every frame with `transforms.Resize((image_size, image_size))` on a PIL image
(core/preprocessing/frame_loader.py:34-45), i.e. Pillow's ImagingResample with the bilinear ("triangle") filter, which
widens the filter support by the scale factor when shrinking (antialiasing) and quantises the normalised weights to
22 fractional bits.  The tables built here are what `vc_resize_bilinear_u8` consumes; the per-pixel work is on the GPU.
Double-precision arithmetic in the same order as Pillow's `precompute_coeffs` / `normalize_coeffs_8bpc`.
"""
from __future__ import annotations

import ctypes as C
from functools import lru_cache

import numpy as np
import torch

from . import lib as L

PRECISION_BITS = 32 - 8 - 2


@lru_cache(maxsize=64)
def pillow_bilinear_coeffs(in_size: int, out_size: int):
    """-> (weights int32 [out, ksize], bounds int32 [out, 2] = (first input index, number of taps), ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale                      # bilinear filter support 1.0
    ksize = int(np.ceil(support)) * 2 + 1
    inv = 1.0 / filterscale
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        k = np.zeros(ksize, dtype=np.float64)
        ww = 0.0
        for x in range(xmax):
            v = abs((x + xmin - center + 0.5) * inv)
            w = 1.0 - v if v < 1.0 else 0.0
            k[x] = w
            ww += w
        if ww != 0.0:
            for x in range(xmax):
                k[x] /= ww
        for x in range(ksize):
            kk[xx, x] = int(-0.5 + k[x] * (1 << PRECISION_BITS)) if k[x] < 0 else int(0.5 + k[x] * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return kk, bounds, ksize


class FrameResizer:
    """uint8 [n,H,W,3] on the device -> uint8 [n,out,out,3], byte-exact with PIL.Image.resize(BILINEAR).  Coefficient tables
    and scratch buffers are cached per input size; everything is enqueued on the current stream."""

    def __init__(self, device, out_h: int, out_w: int):
        self.device = torch.device(device)
        self.out_h, self.out_w = int(out_h), int(out_w)
        self._tables: dict = {}
        self._scratch = None

    def _table(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        t = self._tables.get(key)
        if t is None:
            kk, b, ks = pillow_bilinear_coeffs(in_size, out_size)
            t = (torch.from_numpy(kk).to(self.device), torch.from_numpy(b).to(self.device), ks)
            self._tables[key] = t
        return t

    def __call__(self, frames_u8: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        if frames_u8.dtype != torch.uint8 or frames_u8.ndim < 3 or frames_u8.shape[-1] != 3:
            raise ValueError(f"expect uint8 [...,H,W,3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
        if frames_u8.device != self.device:
            raise ValueError("frames must live on the resizer's device")
        lead = frames_u8.shape[:-3]
        H, W = int(frames_u8.shape[-3]), int(frames_u8.shape[-2])
        src = frames_u8.contiguous().view(-1, H, W, 3)
        n = src.shape[0]
        if out is None:
            out = torch.empty(n, self.out_h, self.out_w, 3, dtype=torch.uint8, device=self.device)
        if n == 0:
            return out.view(*lead, self.out_h, self.out_w, 3)
        kx, bx, ksx = self._table(W, self.out_w)
        ky, by, ksy = self._table(H, self.out_h)
        need = n * H * self.out_w * 3
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        lib = L.load()
        L.check(lib.vc_resize_bilinear_u8(src.data_ptr(), n, H, W, self._scratch.data_ptr(), out.data_ptr(), self.out_h, self.out_w,
                                          kx.data_ptr(), bx.data_ptr(), ksx, ky.data_ptr(), by.data_ptr(), ksy, L.current_stream(self.device)))
        return out.view(*lead, self.out_h, self.out_w, 3)
