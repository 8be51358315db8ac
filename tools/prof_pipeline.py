"""Steady-state pipeline throughput vs decode length and decode grouping: how much of the batch time the decode chain costs."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel

a = synthetic.ARCHS["vit_b16_gpt2"]
B, T = 64, 16
dev = torch.device("cuda", 0)
m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), dev, vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, chunk_frames=1024)
devf = synthetic.make_batch_u8(0, B, T).to(dev)


def run(pipe, steps):
    for _ in range(6):
        pipe.submit(devf, to_host=False)
    pipe.drain(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in (pipe.copy_stream, pipe.enc_stream, pipe.dec_stream):
        s_.wait_event(e0)
    for i in range(steps):
        pipe.submit(devf, to_host=False)
    torch.cuda.current_stream().wait_event(pipe.last_event())
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


import os
cfgs = [(20, 2, True), (20, 2, False), (20, 4, True), (20, 4, False), (20, 1, False)]
for rep in range(2):
    for n_new, group, ov in cfgs:
        pipe = m.pipeline(max_new_tokens=n_new, decode_group=group, overlap_decode=ov)
        ms = run(pipe, 24)
        print(f"max_new={n_new:2d} group={group} overlap={ov}: {ms:6.2f} ms/batch  {64 / ms * 1e3:7.1f} captions/s")
# encoder alone, back to back on one stream
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    m.encode_prefix(devf)
e0.record()
for _ in range(24):
    m.encode_prefix(devf)
e1.record(); torch.cuda.synchronize()
print(f"encode_prefix alone, back to back: {e0.elapsed_time(e1) / 24:6.2f} ms/batch")
