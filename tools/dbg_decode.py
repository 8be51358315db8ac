"""A/B: persistent decode kernel vs the per-op fallback (VC_DECODE_FALLBACK=1 in a second process)."""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel

arch = sys.argv[1] if len(sys.argv) > 1 else "tiny"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
n_new = 6
a = synthetic.ARCHS[arch]
sd = synthetic.make_state_dict(a, seed=1234)
m = B200CaptionModel(sd, "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads)
g = torch.Generator().manual_seed(5)
prefix = torch.randn(B, a.prefix_len, a.gpt_dim, generator=g).cuda() * 0.3
forced = torch.randint(0, 50000, (B, n_new), generator=g).cuda().int()
ids, lens, logits = m.greedy_ids(prefix, None, n_new, forced_ids=forced, keep_logits=True, use_graph=False)
torch.cuda.synchronize()
out = Path("gpurun_out") / ("dbg_logits_%s.pt" % ("fallback" if os.environ.get("VC_DECODE_FALLBACK") else "kernel"))
torch.save(dict(logits=logits.cpu().clone(), ids=ids.cpu().clone()), out)
other = Path("gpurun_out") / "dbg_logits_fallback.pt"
if not os.environ.get("VC_DECODE_FALLBACK") and other.exists():
    ref = torch.load(other)
    for s in range(n_new):
        d = (logits[s].cpu() - ref["logits"][s]).abs()
        print(f"step {s}: max abs diff {d.max().item():.4f}  mean {d.mean().item():.5f}  argmax eq {(logits[s].cpu().argmax(-1) == ref['logits'][s].argmax(-1)).float().mean().item():.2f}")
    print("ids kernel", ids.cpu().tolist(), "fallback", ref["ids"].tolist())
