"""Beam-search timing for BASELINE.json configs[3] (64 videos, 5 beams, 30 tokens): per-kernel CUDA-event spans of one eager
search and the graph-replay time of the whole search.   python tools/prof_beam.py [videos [beams [tokens]]]"""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa: F401
from vcb200 import lib as L, synthetic
from vcb200.beam import beam_search_ids
from vcb200.model import B200CaptionModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 5
n_new = int(sys.argv[3]) if len(sys.argv) > 3 else 30
a = synthetic.ARCHS["vit_b16_gpt2"]
m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads)
lib = L.load()
prefix = torch.randn(B, a.prefix_len, a.gpt_dim, device="cuda") * 0.1


def run(graph):
    with torch.cuda.device(0):
        return beam_search_ids(m, prefix, [50256], max_new_tokens=n_new, num_beams=nb, use_graph=graph)


for _ in range(2):
    run(False)
torch.cuda.synchronize()
lib.vc_prof_begin()
run(False)
mx = 64
names = C.create_string_buffer(mx * 48); tms = (C.c_float * mx)(); calls = (C.c_int * mx)(); work = (C.c_double * mx)()
n = lib.vc_prof_end(mx, names, tms, calls, work)
tot = 0
print(f"--- {B} videos x {nb} beams, {n_new} tokens: eager launches, event span per kernel")
for i in range(n):
    nm = names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode()
    print(f"{nm:20s} calls {calls[i]:5d} total {tms[i]:8.3f} ms  avg {tms[i] / calls[i] * 1e3:8.2f} us")
    tot += tms[i]
print("sum of kernel event spans", round(tot, 3), "ms")
for _ in range(3):
    run(True)
ts = []
for _ in range(20):
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(); run(True); t1.record(); torch.cuda.synchronize()
    ts.append(t0.elapsed_time(t1))
ts.sort()
print(f"graph replay of the whole search: p50 {ts[len(ts) // 2]:.3f} ms ({ts[len(ts) // 2] / n_new * 1e3:.1f} us per token), min {ts[0]:.3f} ms")
