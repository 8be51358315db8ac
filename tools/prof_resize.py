"""Times the Pillow-exact resize on MSVD-like frame sizes (1024 frames) and reports GB/s of algorithmic bytes."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200.resample import FrameResizer

rs = FrameResizer("cuda:0", 224, 224)
for (H, W) in [(240, 320), (360, 480), (720, 1280)]:
    n = 1024 if H < 700 else 256
    x = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty(n, 224, 224, 3, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        rs(x, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        rs(x, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    alg = n * 3 * (H * W + 2 * H * 224 + 224 * 224)       # read input, write + read the uint8 intermediate, write output
    print(f"{n} frames {H}x{W} -> 224x224: {ms:.3f} ms  {alg / ms / 1e6:.0f} GB/s algorithmic ({ms / n * 1e3:.2f} us/frame)")

# per-kernel split (CUDA events around each launch)
import ctypes as C
from vcb200 import lib as L
lib = L.load()
x = torch.randint(0, 256, (1024, 360, 480, 3), dtype=torch.uint8, device="cuda")
out = torch.empty(1024, 224, 224, 3, dtype=torch.uint8, device="cuda")
rs(x, out); torch.cuda.synchronize()
lib.vc_prof_begin()
for _ in range(5):
    rs(x, out)
mx = 16
names = C.create_string_buffer(mx * 48); tms = (C.c_float * mx)(); calls = (C.c_int * mx)(); work = (C.c_double * mx)()
n = lib.vc_prof_end(mx, names, tms, calls, work)
for i in range(n):
    nm = names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode()
    print(f"{nm:12s} {tms[i] / calls[i]:.3f} ms per launch, {work[i] / calls[i] / (tms[i] / calls[i]) / 1e6:.0f} GB/s of its own read+write bytes")
