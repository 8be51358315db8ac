"""Per-kernel CUDA-event breakdown of one encoder pass (eager launches) for a named architecture."""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import lib as L, synthetic
from vcb200.model import B200CaptionModel

arch = sys.argv[1] if len(sys.argv) > 1 else "vit_l14_gpt2m"
B, T = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16, 32)
a = synthetic.ARCHS[arch]
m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, chunk_frames=1024)
frames = synthetic.make_batch_u8(0, B, T).cuda()
lib = L.load()
for _ in range(2):
    m.encode_prefix(frames)
torch.cuda.synchronize()
lib.vc_prof_begin()
m.encode_prefix(frames)
mx = 64
names = C.create_string_buffer(mx * 48); tms = (C.c_float * mx)(); calls = (C.c_int * mx)(); work = (C.c_double * mx)()
n = lib.vc_prof_end(mx, names, tms, calls, work)
tot = sum(tms[i] for i in range(n))
for i in range(n):
    nm = names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode()
    extra = f"  {work[i] / tms[i] / 1e9:8.1f} TFLOP/s" if nm.startswith("gemm") else ""
    print(f"{nm:20s} calls {calls[i]:4d}  {tms[i]:8.3f} ms  {100 * tms[i] / tot:5.1f} %{extra}")
print(f"sum {tot:.3f} ms for {B * T} frames")
