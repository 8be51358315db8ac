"""Encoder pass (64 videos x 16 frames) against the number of frames per encoder chunk: with small chunks the activations of
consecutive kernels stay in the 126 MB L2 (less HBM traffic, less power), at the price of more and shorter launches.
usage: sweep_chunk_frames.py [chunk ...]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel

chunks = [int(x) for x in sys.argv[1:]] or [1024, 256, 128, 64, 32]
a = synthetic.ARCHS["vit_b16_gpt2"]
sd = synthetic.make_state_dict(a, seed=1234)
frames = synthetic.make_batch_u8(0, 64, 16).cuda()
models = {c: B200CaptionModel(sd, "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, chunk_frames=c) for c in chunks}
for c in chunks:
    for _ in range(2):
        models[c].encode_prefix(frames)
torch.cuda.synchronize()
res = {c: [] for c in chunks}
for r in range(3):
    for c in chunks:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            models[c].encode_prefix(frames)
        e1.record(); torch.cuda.synchronize()
        res[c].append(e0.elapsed_time(e1) / 5)
for c in chunks:
    print(f"chunk_frames {c:5d}: " + "  ".join(f"{x:6.2f}" for x in res[c]) + f"   mean {sum(res[c]) / len(res[c]):6.2f} ms per 1024 frames")
