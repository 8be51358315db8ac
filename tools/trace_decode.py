"""In-kernel timeline of the decode chain (globaltimer stamps written by the kernels of csrc/decode_chain.cu): where a decode step's
time goes between kernel entry, the dependency wait, operand arrival, MMAs and the reduction.  python tools/trace_decode.py [n_seq]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa: F401
from vcb200 import lib as L, synthetic
from vcb200.model import B200CaptionModel

a = synthetic.ARCHS["vit_b16_gpt2"]
sd = synthetic.make_state_dict(a, seed=1234)
m = B200CaptionModel(sd, "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads)
lib = L.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_new = 4
prefix = torch.randn(B, a.prefix_len, a.gpt_dim, device="cuda") * 0.1
for _ in range(3):
    m.greedy_ids(prefix, None, n_new)          # graph captured and replayed
torch.cuda.synchronize()
MAXR = 4096
buf = torch.zeros(8 + 8 * MAXR, dtype=torch.int64, device="cuda")
L.check(lib.vc_debug_trace(buf.data_ptr(), MAXR))
m.greedy_ids(prefix, None, n_new)
torch.cuda.synchronize()
L.check(lib.vc_debug_trace(0, 0))
n = int(buf[0].item())
rec = buf[8:8 + 8 * n].view(n, 8).cpu().tolist()
names = {1: "add_pos", 2: "qkv", 3: "proj", 4: "fc1", 5: "fc2", 6: "lm_head", 7: "select"}
rec.sort(key=lambda r: r[1])
t0 = rec[0][1]
print(f"{n} records; columns: entry, +wait_done, +x_arrived, +mma_done, +reduce_barrier, exit (us relative to first entry / to entry)")
last_exit = None
for r in rec:
    kid, cta = r[0] & 0xff, r[0] >> 8
    e = r[1]
    pts = [(r[i] - e) / 1e3 if r[i] else None for i in (2, 3, 4, 5, 7)]
    s = " ".join(f"{p:7.2f}" if p is not None else "      -" for p in pts)
    print(f"{names.get(kid, kid):8s} cta {cta:4d} entry {(e - t0) / 1e3:9.2f}  {s}")
