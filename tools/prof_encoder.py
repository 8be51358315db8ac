"""One preprocess + ViT encode + prefix pass over 64 x 16 frames (target of the ncu launch list / DRAM-traffic pass)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel
a = synthetic.ARCHS["vit_b16_gpt2"]
sd = synthetic.make_state_dict(a, seed=1234)
m = B200CaptionModel(sd, "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, chunk_frames=1024)
frames = synthetic.make_batch_u8(0, 64, 16).cuda()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(n):
    m.encode_prefix(frames)
torch.cuda.synchronize()
print("ok")
