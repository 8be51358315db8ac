"""Per-kernel CUDA-event timing of the greedy decode (eager launches) for the bench workload."""
import ctypes as C
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa: F401
from vcb200 import lib as L, synthetic
from vcb200.model import B200CaptionModel

a = synthetic.ARCHS["vit_b16_gpt2"]
sd = synthetic.make_state_dict(a, seed=1234)
m = B200CaptionModel(sd, "cuda:0", vit_heads=12, gpt_heads=12)
lib = L.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
prefix = torch.randn(B, 4, 768, device="cuda") * 0.1
for _ in range(2):
    m.greedy_ids(prefix, None, 20, use_graph=False)
torch.cuda.synchronize()
lib.vc_prof_begin()
m.greedy_ids(prefix, None, 20, use_graph=False)
mx = 64
names = C.create_string_buffer(mx * 48); tms = (C.c_float * mx)(); calls = (C.c_int * mx)(); work = (C.c_double * mx)()
n = lib.vc_prof_end(mx, names, tms, calls, work)
tot = 0
for i in range(n):
    nm = names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode()
    print(f"{nm:20s} calls {calls[i]:5d} total {tms[i]:8.3f} ms  avg {tms[i] / calls[i] * 1e3:8.2f} us")
    tot += tms[i]
print("sum of kernel event spans", round(tot, 3), "ms")
for use_graph in (True,):
    m.greedy_ids(prefix, None, 20); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        m.greedy_ids(prefix, None, 20)
    e1.record(); torch.cuda.synchronize()
    print("graph replay greedy x20 tokens:", e0.elapsed_time(e1) / 5, "ms")
