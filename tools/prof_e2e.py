"""Resident vs host-buffer (pinned H2D + ids D2H) pipeline throughput in alternating blocks: separates the cost of the
copies from the power-cap clock drift between two back-to-back measurements."""
import subprocess, sys
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
host = synthetic.make_batch_u8(0, B, T).pin_memory()
devf = host.to(dev)
pipe = m.pipeline(max_new_tokens=20, decode_group=4)
pipe.warm(devf)


def clock():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()


def block(fn, steps):
    pipe.drain(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in (pipe.copy_stream, pipe.enc_stream, pipe.dec_stream):
        s_.wait_event(e0)
    for i in range(steps):
        fn()
    torch.cuda.current_stream().wait_event(pipe.last_event())
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, ''


res = lambda: pipe.submit(devf, to_host=False)
e2e = lambda: pipe.submit(host, to_host=True)
for _ in range(4):
    res()
for name, fn in [("resident", res), ("e2e", e2e)] * 4:
    ms, c = block(fn, 20)
    print(f"{name:9s} {ms:7.2f} ms/batch  {64 / ms * 1e3:7.1f} captions/s   [{c}]")
# the copy alone
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dst = torch.empty_like(devf)
e0.record()
for _ in range(10):
    dst.copy_(host, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"H2D of {host.numel() / 1e6:.0f} MB alone: {e0.elapsed_time(e1) / 10:.2f} ms ({host.numel() / e0.elapsed_time(e1) * 10 / 1e6:.1f} GB/s)")
