"""JPEG frame loading off the single host thread (second half).

The reference decodes every frame with PIL on ONE thread inside the timed window (core/preprocessing/frame_loader.py:42-45:
`Image.open(p).convert("RGB")` per file).  Here the same decoder — so the bytes are identical by construction — runs in a pool of
host threads (Pillow releases the GIL while it decodes) and writes straight into a pinned uint8 [B,T,H,W,3] buffer that the
pipeline copies to the device asynchronously; the resize to 224 x 224 (frame_loader.py:36) then runs on the GPU, byte-exact with
Pillow's (`vc_resize_bilinear_u8`).  nvJPEG was considered and not used: its IDCT / chroma upsampling are not bit-identical to
libjpeg-turbo's, and this path's contract for preprocessing is bit-exactness.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import List, Sequence, Union

import numpy as np
import torch

from .engine import sample_frame_indices

Source = Union[str, os.PathLike, bytes]


def list_sampled_frames(frames_dir: Union[str, os.PathLike], num_frames: int) -> List[Path]:
    """frame_loader.py:13-16, :31-32: sorted frame_*.jpg, step = max(n // T, 1), files[::step][:T] (no padding)."""
    files = sorted(Path(frames_dir).glob("frame_*.jpg"))
    if not files:
        raise FileNotFoundError(f"No frame_*.jpg files found under {frames_dir}")
    return [files[i] for i in sample_frame_indices(len(files), num_frames)]


def _decode_into(src: Source, dst: np.ndarray) -> None:
    import io
    from PIL import Image
    with Image.open(io.BytesIO(src) if isinstance(src, (bytes, bytearray)) else src) as im:
        a = np.asarray(im.convert("RGB"))
    if a.shape != dst.shape:
        raise ValueError(f"frame size {a.shape[:2]} differs from {dst.shape[:2]} within one batch")
    dst[...] = a


class FrameDecodePool:
    """Decodes batches of JPEG frames (paths or encoded bytes) into pinned host memory with `workers` threads."""

    def __init__(self, workers: int = 0, pin: bool = True):
        self.workers = int(workers) if workers and workers > 0 else min(32, os.cpu_count() or 1)
        self._pool = ThreadPoolExecutor(max_workers=self.workers, thread_name_prefix="vcb200-jpeg")
        self._pin = bool(pin) and torch.cuda.is_available()
        self._buf: dict = {}

    def close(self) -> None:
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _buffer(self, shape) -> torch.Tensor:
        t = self._buf.get(shape)
        if t is None:
            t = torch.empty(shape, dtype=torch.uint8)
            if self._pin:
                t = t.pin_memory()
            self._buf = {shape: t}                 # one staging buffer (the latest shape): bounded page-locked memory
        return t

    def decode(self, videos: Sequence[Sequence[Source]]) -> torch.Tensor:
        """videos[b][t] = path or JPEG bytes of frame t of video b (same T and frame size for the whole batch).
        Returns uint8 [B,T,H,W,3] (pinned when CUDA is present); valid until the next call."""
        B = len(videos)
        if B == 0:
            return torch.empty(0, 0, 0, 0, 3, dtype=torch.uint8)
        T = len(videos[0])
        if any(len(v) != T for v in videos):
            raise ValueError("every video of a batch must have the same number of sampled frames")
        import io
        from PIL import Image
        first = videos[0][0]
        with Image.open(io.BytesIO(first) if isinstance(first, (bytes, bytearray)) else first) as im:
            W, H = im.size
        out = self._buffer((B, T, H, W, 3))
        arr = out.numpy()
        futs = [self._pool.submit(_decode_into, videos[b][t], arr[b, t]) for b in range(B) for t in range(T)]
        for f in futs:
            f.result()
        return out

    def decode_dirs(self, frames_dirs: Sequence[Union[str, os.PathLike]], num_frames: int) -> torch.Tensor:
        """One clip directory per video, sampled like the reference's load_video_tensor."""
        return self.decode([list_sampled_frames(d, num_frames) for d in frames_dirs])
