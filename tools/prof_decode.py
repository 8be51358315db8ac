"""Greedy-decode timing for the bench workload: per-kernel CUDA-event spans (eager launches) and the step latency from graph
replays, (T(prefill + n-1 steps) - T(prefill)) / (n-1).   python tools/prof_decode.py [n_seq ...]   (VC_DECODE_CHAIN=1: round-1 chain)"""
import ctypes as C
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa: F401
from vcb200 import lib as L, synthetic
from vcb200.model import B200CaptionModel

arch = os.environ.get("VC_ARCH", "vit_b16_gpt2")
a = synthetic.ARCHS[arch]
sd = synthetic.make_state_dict(a, seed=1234)
m = B200CaptionModel(sd, "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads)
lib = L.load()
sizes = [int(x) for x in sys.argv[1:]] or [64]
n_new = 20
for B in sizes:
    prefix = torch.randn(B, a.prefix_len, a.gpt_dim, device="cuda") * 0.1
    for _ in range(2):
        m.greedy_ids(prefix, None, n_new, use_graph=False)
    torch.cuda.synchronize()
    lib.vc_prof_begin()
    m.greedy_ids(prefix, None, n_new, use_graph=False)
    mx = 64
    names = C.create_string_buffer(mx * 48); tms = (C.c_float * mx)(); calls = (C.c_int * mx)(); work = (C.c_double * mx)()
    n = lib.vc_prof_end(mx, names, tms, calls, work)
    tot = 0
    print(f"--- n_seq {B} ({arch}, chain {'v1' if os.environ.get('VC_DECODE_CHAIN') == '1' else 'v2'}) eager launches, event span per kernel")
    for i in range(n):
        nm = names.raw[i * 48:(i + 1) * 48].split(b"\0")[0].decode()
        print(f"{nm:20s} calls {calls[i]:5d} total {tms[i]:8.3f} ms  avg {tms[i] / calls[i] * 1e3:8.2f} us")
        tot += tms[i]
    print("sum of kernel event spans", round(tot, 3), "ms")
    m.greedy_ids(prefix, None, n_new); m.greedy_ids(prefix, None, 1); torch.cuda.synchronize()
    full, pre = [], []
    for _ in range(5):
        m.greedy_ids(prefix, None, n_new); m.greedy_ids(prefix, None, 1)
    for _ in range(30):
        t0, t1, t2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        t0.record(); m.greedy_ids(prefix, None, n_new); t1.record(); m.greedy_ids(prefix, None, 1); t2.record()
        torch.cuda.synchronize()
        full.append(t0.elapsed_time(t1)); pre.append(t1.elapsed_time(t2))
    step = sorted((f - p) / (n_new - 1) * 1e3 for f, p in zip(full, pre))
    print(f"n_seq {B}: graph replay {n_new} tokens p50 {sorted(full)[len(full) // 2]:.3f} ms, prefill+1 {sorted(pre)[len(pre) // 2]:.3f} ms, "
          f"step p50 {step[len(step) // 2]:.1f} us (min {step[0]:.1f})")
