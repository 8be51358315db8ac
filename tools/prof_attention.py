"""Times vit_attention on ViT-B/16 shapes (target of ncu -k regex:vit_attention)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import lib as L
lib = L.load()
frames, tokens, heads = int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 197, 12
D = heads * 64
qkv = torch.randn(frames * tokens, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.zeros(frames * tokens, D, device="cuda", dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    L.check(lib.vc_vit_attention(qkv.data_ptr(), out.data_ptr(), frames, tokens, heads, 64, st))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    L.check(lib.vc_vit_attention(qkv.data_ptr(), out.data_ptr(), frames, tokens, heads, 64, st))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
fl = 4.0 * frames * heads * tokens * tokens * 64
print(f"vit_attention frames={frames}: {ms:.4f} ms  {fl / ms / 1e9:.1f} TFLOP/s  ({ms * 1e3 * 148 / (frames * heads):.2f} us*SM per (frame,head))")
