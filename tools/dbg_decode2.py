"""Stage-by-stage check of the persistent decode kernel on a 1-layer GPT-2 (buffers left in the workspace after one step)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, math
import torch.nn.functional as F
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
a0 = synthetic.ARCHS["tiny"]
a = synthetic.Arch("one", a0.image, a0.patch, a0.vit_dim, 1, a0.vit_heads, a0.vit_mlp, a0.video_dim, a0.gpt_dim, 1, a0.gpt_heads, a0.vocab, a0.n_pos, a0.prefix_len) if hasattr(a0, "image") else None
if a is None:
    import dataclasses
    a = dataclasses.replace(a0, name="one", vit_layers=1, gpt_layers=1)
sd = synthetic.make_state_dict(a, seed=1234)
m = B200CaptionModel(sd, "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads)
H, heads = a.gpt_dim, a.gpt_heads
g = torch.Generator().manual_seed(7)
x0 = (torch.randn(B, 5, H, generator=g) * 0.3).cuda()
x1 = (torch.randn(B, 1, H, generator=g) * 0.3).cuda()
gpt2 = m.decoder.model
out0 = gpt2(inputs_embeds=x0, past_key_values=None, use_cache=True, return_dict=True, s_max=32)
cache = out0.past_key_values
out1 = gpt2(inputs_embeds=x1, past_key_values=cache, use_cache=True, return_dict=True)
torch.cuda.synchronize()
ws = m.ws.gpt(B, B)
al = lambda v: (v + 1023) // 1024 * 1024
R = B
off = 0
def take(nbytes, dtype, shape):
    global off
    t = ws[off:off + nbytes].view(dtype).view(*shape).clone()
    off += al(nbytes)
    return t
h = take(R * H * 4, torch.float32, (R, H))
xn = take(R * H * 2, torch.bfloat16, (R, H))
qkv = take(R * 3 * H * 2, torch.bfloat16, (R, 3 * H))
att = take(R * H * 2, torch.bfloat16, (R, H))
hid = take(R * 4 * H * 2, torch.bfloat16, (R, 4 * H))
emb_ = take(R * H * 4, torch.float32, (R, H))
logits_ = take(B * m.dims["vocab_pad"] * 4, torch.float32, (B, m.dims["vocab_pad"]))
fin_ = take(B * 4, torch.int32, (B,))
nxt_ = take(B * 4, torch.int32, (B,))
part = take(9216 * B * 4, torch.float32, (9216 * B,))
import os
STOP = int(os.environ.get("VC_DK_STOP", "0"))

# fp32 reference on the GPU, weights rounded to bf16 like the packed model
p = "decoder.model.transformer."
W = lambda k: sd[p + k].cuda().float()
bf = lambda t: t.to(torch.bfloat16).float()
def ln(x, w, b): return F.layer_norm(x, (H,), w, b, 1e-5)
def block_inputs(x, past_len):
    return x + W("wpe.weight")[past_len:past_len + x.shape[1]]
def attn_kv(xn_):
    qkv_ = bf(xn_) @ bf(W("h.0.attn.c_attn.weight")) + W("h.0.attn.c_attn.bias")
    return bf(qkv_)
hp = block_inputs(x0, 0)
qkv_p = attn_kv(ln(hp, W("h.0.ln_1.weight"), W("h.0.ln_1.bias")))
k_p, v_p = qkv_p[..., H:2 * H], qkv_p[..., 2 * H:]
h1 = block_inputs(x1, 5)
qkv1 = attn_kv(ln(h1, W("h.0.ln_1.weight"), W("h.0.ln_1.bias")))
xn1 = ln(h1, W("h.0.ln_1.weight"), W("h.0.ln_1.bias"))[:, 0]
if STOP in (1, 2, 3):
    print("h(embed) err", (h - h1[:, 0]).abs().max().item())
    print("xn(ln1) err", (xn.float() - xn1).abs().max().item())
if STOP in (2, 3):
    P = part[: 4 * B * 3 * H].view(4, B, 3 * H)
    full = bf(xn1) @ bf(W("h.0.attn.c_attn.weight"))          # [B,3H] without bias
    got = P.sum(0)
    err = (got - full).abs()
    print("qkv partial-sum err max", err.max().item(), " per 64-col tile max:", [round(v, 3) for v in err.view(B, 36, 64).amax(dim=(0, 2)).tolist()])
    for ksl in range(4):
        ref_k = bf(xn1)[:, ksl * 192:(ksl + 1) * 192] @ bf(W("h.0.attn.c_attn.weight"))[ksl * 192:(ksl + 1) * 192]
        e = (P[ksl] - ref_k).abs().view(B, 36, 64).amax(dim=(0, 2))
        print("  kslice", ksl, "tiles with err>0.01:", [i for i, v in enumerate(e.tolist()) if not (v < 0.01)])
    if STOP == 2: sys.exit(0)
kv = cache.kv  # [layers,2,n_seq,heads,s_max,64]
k_new = kv[0, 0, :, :, 5, :].reshape(B, H).float()
v_new = kv[0, 1, :, :, 5, :].reshape(B, H).float()
ek = (k_new - qkv1[:, 0, H:2 * H]).abs().view(B, heads, 64)
print("k_new err per (seq,head):", [[("%.2g" % v) for v in row] for row in ek.amax(-1).tolist()])
print("k_new sample", k_new[0, :8].tolist(), "ref", qkv1[0, 0, H:H + 8].tolist())
ea = None
print("k_new err", (k_new - qkv1[:, 0, H:2 * H]).abs().max().item(), " v_new err", (v_new - qkv1[:, 0, 2 * H:]).abs().max().item(), " (scale", qkv1.abs().max().item(), ")")
k_old = kv[0, 0, :, :, :5, :].float()   # [B,heads,5,64]
print("k_old (prefill) err", (k_old - k_p.view(B, 5, heads, 64).permute(0, 2, 1, 3)).abs().max().item())
q = qkv1[:, 0, :H].view(B, heads, 1, 64)
K = torch.cat([k_p, qkv1[..., H:2 * H]], 1).view(B, 6, heads, 64).permute(0, 2, 1, 3)
V = torch.cat([v_p, qkv1[..., 2 * H:]], 1).view(B, 6, heads, 64).permute(0, 2, 1, 3)
pr = torch.softmax(q @ K.transpose(-1, -2) / 8.0, -1)
att_ref = (pr @ V).permute(0, 2, 1, 3).reshape(B, H)
print("att err", (att.float() - att_ref).abs().max().item(), " (scale", att_ref.abs().max().item(), ")")
print("att err per (seq,head):", [[("%.2g" % v) for v in row] for row in (att.float() - att_ref).abs().view(B, heads, 64).amax(-1).tolist()])
if STOP == 3: sys.exit(0)
ap = bf(att_ref) @ bf(W("h.0.attn.c_proj.weight")) + W("h.0.attn.c_proj.bias")
h2 = h1[:, 0] + ap
xn2 = ln(h2, W("h.0.ln_2.weight"), W("h.0.ln_2.bias"))
if STOP == 4:
    P = part[: 6 * B * H].view(6, B, H)
    e = (P.sum(0) - (ap - W("h.0.attn.c_proj.bias"))).abs()
    print("aproj partial-sum err", e.max().item(), "per 32-col tile:", [round(v, 3) for v in e.view(B, 24, 32).amax(dim=(0, 2)).tolist()])
    sys.exit(0)
if STOP == 5:
    print("h(after attn) err", (h - h2).abs().max().item(), " xn(ln2) err", (xn.float() - xn2).abs().max().item())
    sys.exit(0)
fc = bf(xn2) @ bf(W("h.0.mlp.c_fc.weight")) + W("h.0.mlp.c_fc.bias")
hid_ref = F.gelu(fc, approximate="tanh")
print("hid err", (hid.float() - hid_ref).abs().max().item(), " (scale", hid_ref.abs().max().item(), ")")
hid_stale = F.gelu(bf(xn1) @ bf(W("h.0.mlp.c_fc.weight")) + W("h.0.mlp.c_fc.bias"), approximate="tanh")
print("hid vs STALE-xn (LN1 output) reference:", (hid.float() - hid_stale).abs().max().item())
eh = (hid.float() - hid_ref).abs().view(B, 96, 32).amax(dim=(0, 2))
print("hid bad 32-col tiles:", [i for i, v in enumerate(eh.tolist()) if not (v < 0.02)])
if STOP == 6: sys.exit(0)
mp = bf(hid_ref) @ bf(W("h.0.mlp.c_proj.weight")) + W("h.0.mlp.c_proj.bias")
h3 = h2 + mp
print("h err", (h - h3).abs().max().item(), " (scale", h3.abs().max().item(), ")")
xnf = ln(h3, W("ln_f.weight"), W("ln_f.bias"))
print("xn err", (xn.float() - xnf).abs().max().item())
lg = bf(xnf) @ bf(W("wte.weight")).t()
print("logits err", (out1.logits[:, 0].float() - lg).abs().max().item(), " (scale", lg.abs().max().item(), ")")
