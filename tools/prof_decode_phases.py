"""Per-phase timeline of the persistent decode kernel (VC_DK_PROF=1): CTA 0 stamps %globaltimer when it observes each
phase complete.  Prints the mean duration of each phase kind over the steps of one greedy run."""
import os, sys
os.environ["VC_DK_PROF"] = "1"
os.environ["VC_DECODE_PERSISTENT"] = sys.argv[3] if len(sys.argv) > 3 else "1"
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel

arch = sys.argv[1] if len(sys.argv) > 1 else "vit_b16_gpt2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
n_new = 8
a = synthetic.ARCHS[arch]
sd = synthetic.make_state_dict(a, seed=1234)
m = B200CaptionModel(sd, "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads)
g = torch.Generator().manual_seed(5)
prefix = (torch.randn(B, a.prefix_len, a.gpt_dim, generator=g) * 0.3).cuda()
for _ in range(3):
    m.greedy_ids(prefix, None, n_new, use_graph=False)
torch.cuda.synchronize()
st = next(v for k, v in m._graphs.items() if k[0] == "greedy")
ws = st["ws"]
H, L, V = a.gpt_dim, a.gpt_layers, m.dims["vocab_pad"]
al = lambda v: (v + 1023) // 1024 * 1024
R = B * (a.prefix_len + 1)
off = al(R * H * 4) + al(R * H * 2) + al(R * 3 * H * 2) + al(R * H * 2) + al(R * 4 * H * 2) + al(R * H * 4) + al(B * V * 4) + al(B * 4) + al(B * 4)
rows = min(B, 128)
part_floats = max(4 * 3 * H, 6 * H, 12 * H) if H == 768 else None
import ctypes as C
from vcb200 import lib as Lb
# partial size: ask the library indirectly through the total size
total = Lb.load().vc_gpt_workspace_bytes(C.byref(m.packed.gpt), B, R)
cand_i_off = total - 1024 - 256 * 128 * 4
prof = ws[cand_i_off + 148 * 128 * 4: cand_i_off + 256 * 128 * 4].view(torch.int64).cpu()
per_step = 7 * L + 2
n_steps = n_new - 1
names = ["qkv", "att", "aproj", "ln2", "fc1", "fc2", "ln1"]
t = prof[: 1 + n_steps * per_step].tolist()
# t[k] = time CTA 0 finished phase k (and had seen phase k-1 complete everywhere); duration of phase k = t[k]-t[k-1]
import collections
acc = collections.defaultdict(list)
for s in range(1, n_steps):          # skip the first step (cold)
    base = 1 + s * per_step
    for l in range(L):
        for i, nm in enumerate(names):
            k = base + 7 * l + i
            acc[nm].append((t[k] - t[k - 1]) / 1e3)
    acc["lm_head"].append((t[base + 7 * L] - t[base + 7 * L - 1]) / 1e3)
    acc["select"].append((t[base + 7 * L + 1] - t[base + 7 * L]) / 1e3)
    acc["step"].append((t[base + per_step - 1] - t[base - 1]) / 1e3)
for k, v in acc.items():
    v = [x for x in v if x == x]
    if v:
        print(f"{k:8s} mean {sum(v) / len(v):8.2f} us   min {min(v):8.2f}  max {max(v):8.2f}  n={len(v)}")

if os.environ["VC_DECODE_PERSISTENT"] == "2":
    c0, t0, c1, t1 = prof[2990:2994].tolist()
    print(f"SM clock during the kernel: {(c1 - c0) / (t1 - t0) * 1e3:.0f} MHz over {(t1 - t0) / 1e3:.0f} us")
    st = prof[3000:3000 + 800].tolist()
    tags = {1: "arrive", 2: "released", 10: "gemm in", 11: "x staged", 12: "mma done", 13: "loads issued", 14: "cp.async done", 15: "warp mma", 16: "warp sts", 20: "ln in", 21: "ln summed", 22: "ln done",
            30: "att in", 31: "att qkv", 32: "att scores", 33: "att pv"}
    prev = None
    for i in range(0, 800, 2):
        tag, tm = st[i], st[i + 1]
        if tag == 0 or tag not in tags:
            break
        print(f"  {str(tags.get(tag, tag)):14s} +{(tm - prev) / 1e3 if prev else 0:6.2f} us")
        prev = tm
