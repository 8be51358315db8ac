"""Golden vectors for the frame resize: the reference's OWN transform (frame_loader.py:34-40: torchvision
`transforms.Resize((224, 224))` on a PIL RGB image) applied to structured synthetic frames of several sizes, written to
tests/golden/resize.npz together with the inputs.  Run in the build container (needs Pillow + torchvision):

    python oracle/pin_resize_against_pillow.py
"""
import sys
from pathlib import Path

import numpy as np
import torch
from PIL import Image
from torchvision import transforms

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import vc_oracle as O  # noqa: E402

SIZES = [(120, 160), (37, 53), (224, 300), (100, 90), (225, 223)]     # (H, W): shrink, grow, mixed, near-identity


def frame(h, w, seed):
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    base = np.stack([(xx * 255 // max(w - 1, 1)), (yy * 255 // max(h - 1, 1)), ((xx + yy) * 255 // max(h + w - 2, 1))], axis=-1)
    stripes = ((xx // 3 + yy // 5) % 2 * 90)[..., None]
    noise = g.integers(0, 8, size=(h, w, 1))
    return np.clip(base // 2 + stripes + noise, 0, 255).astype(np.uint8)


def main():
    resize = transforms.Resize((224, 224))
    out = {}
    for i, (h, w) in enumerate(SIZES):
        src = frame(h, w, 100 + i)
        ref = np.asarray(resize(Image.fromarray(src).convert("RGB")))
        mine = O.resize_bilinear_u8(torch.from_numpy(src), 224, 224).numpy()
        assert np.array_equal(ref, mine), (h, w, int(np.abs(ref.astype(int) - mine.astype(int)).max()))
        out[f"src_{h}x{w}"] = src
        out[f"dst_{h}x{w}"] = ref
    # a larger, purely random case checked here but not stored (size)
    for (h, w) in [(360, 480), (480, 640), (720, 1280)]:
        src = np.random.default_rng(h).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        ref = np.asarray(resize(Image.fromarray(src)))
        assert np.array_equal(ref, O.resize_bilinear_u8(torch.from_numpy(src), 224, 224).numpy()), (h, w)
    np.savez_compressed(ROOT / "tests" / "golden" / "resize.npz", **out)
    print("oracle == Pillow on", len(SIZES) + 3, "sizes; wrote tests/golden/resize.npz")


if __name__ == "__main__":
    main()
