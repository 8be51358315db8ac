"""Pipeline settings A/B inside ONE process on ONE box: decode_group x overlap, alternated over several rounds (run-to-run and
box-to-box spread is larger than most effects).  usage: ab_pipeline_modes.py [rounds]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
# optional: the decode groups to compare, e.g. "4,8" (each with the decode chain on its own stream) instead of the default pair
a = synthetic.ARCHS["vit_b16_gpt2"]
dev = torch.device("cuda", 0)
m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), dev, vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, chunk_frames=1024)
devf = synthetic.make_batch_u8(0, 64, 16).to(dev)
MODES = [(int(g), True) for g in sys.argv[2].split(",")] if len(sys.argv) > 2 else [(4, True), (4, False)]
pipes = {}


def run(mode, steps=24):
    if mode not in pipes:
        pipes[mode] = m.pipeline(max_new_tokens=20, decode_group=mode[0], overlap_decode=mode[1])
        pipes[mode].warm(devf)
    pipe = pipes[mode]
    for _ in range(4):
        pipe.submit(devf, to_host=False)
    pipe.drain(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in (pipe.copy_stream, pipe.enc_stream, pipe.dec_stream):
        s_.wait_event(e0)
    for _ in range(steps):
        pipe.submit(devf, to_host=False)
    torch.cuda.current_stream().wait_event(pipe.last_event())
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


res = {mo: [] for mo in MODES}
for r in range(rounds):
    for mo in MODES:
        res[mo].append(run(mo))
for mo in MODES:
    v = res[mo]
    print(f"decode_group={mo[0]} overlap={mo[1]}: " + " ".join(f"{x:6.2f}" for x in v) + f"   mean {sum(v) / len(v):6.2f} ms/batch")
# encoder alone, back to back (what the pipeline could reach if the decode were free)
for _ in range(3):
    m.encode_prefix(devf)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(24):
    m.encode_prefix(devf)
e1.record(); torch.cuda.synchronize()
print(f"encoder alone, back to back: {e0.elapsed_time(e1) / 24:6.2f} ms/batch")
