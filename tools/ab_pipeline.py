"""A/B of library switches inside ONE process on ONE box (run-to-run and box-to-box spread is larger than most effects):
alternates the settings, recapturing the decode graph each time.  usage: ab_pipeline.py ENV_NAME [rounds]"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel

name = sys.argv[1]
for kv in os.environ.get("AB_FIXED", "").split(","):
    if kv:
        os.environ[kv] = "1"
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
a = synthetic.ARCHS["vit_b16_gpt2"]
dev = torch.device("cuda", 0)
m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), dev, vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, chunk_frames=1024)
devf = synthetic.make_batch_u8(0, 64, 16).to(dev)


def run(steps=24):
    m._graphs.clear()
    pipe = m.pipeline(max_new_tokens=20)
    for _ in range(6):
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


res = {0: [], 1: []}
for r in range(rounds):
    for on in (0, 1):
        if on:
            os.environ[name] = "1"
        else:
            os.environ.pop(name, None)
        res[on].append(run())
for on in (0, 1):
    v = res[on]
    print(f"{name}={'1' if on else 'unset'}: " + " ".join(f"{x:6.2f}" for x in v) + f"   mean {sum(v) / len(v):6.2f} ms/batch")
