"""Does the decode chain run underneath the encoder?  Times decode alone, encode alone, and both on two streams."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel

a = synthetic.ARCHS["vit_b16_gpt2"]
sd = synthetic.make_state_dict(a, seed=1234)
m = B200CaptionModel(sd, "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, chunk_frames=1024)
B, T, n_new = 64, 16, 20
frames = synthetic.make_batch_u8(0, B, T).cuda()
feat, prefix = m.encode_prefix(frames)
for _ in range(3):
    m.greedy_ids(prefix, None, n_new)
    m.encode_prefix(frames)
torch.cuda.synchronize()
E, D = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)

def ev():
    return torch.cuda.Event(enable_timing=True)

def run(enc_n, dec_n, label):
    torch.cuda.synchronize()
    e0, e1, d0, d1 = ev(), ev(), ev(), ev()
    with torch.cuda.stream(E):
        e0.record()
        for _ in range(enc_n):
            m.encode_prefix(frames)
        e1.record()
    with torch.cuda.stream(D):
        d0.record()
        for _ in range(dec_n):
            m.greedy_ids(prefix, None, n_new)
        d1.record()
    torch.cuda.synchronize()
    print(f"{label}: encode x{enc_n} {e0.elapsed_time(e1):8.2f} ms   decode x{dec_n} {d0.elapsed_time(d1):8.2f} ms")

run(4, 0, "encode alone")
run(0, 4, "decode alone")
run(4, 4, "both        ")
run(4, 8, "both, 2x dec")

# ---- which encoder kernel slows the decode chain down?  decode x2 under a continuous loop of ONE kernel type
from vcb200 import lib as L
lib = L.load()
M = 1024 * 197
A = (torch.randn(M, 768, device="cuda") * 0.5).to(torch.bfloat16)
W1 = (torch.randn(3072, 768, device="cuda") * 0.05).to(torch.bfloat16)
W0 = (torch.randn(2304, 768, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.randn(3072, device="cuda")
out = torch.zeros(M, 3072, device="cuda", dtype=torch.bfloat16)
x = torch.randn(M, 768, device="cuda")
g = torch.ones(768, device="cuda"); b0 = torch.zeros(768, device="cuda")
xn = torch.empty(M, 768, device="cuda", dtype=torch.bfloat16)
qkv = torch.randn(M, 2304, device="cuda").to(torch.bfloat16)
att = torch.empty(M, 768, device="cuda", dtype=torch.bfloat16)

def load_gemm1(st): L.check(lib.vc_gemm_bf16(A.data_ptr(), W1.data_ptr(), bias.data_ptr(), M, 3072, 768, 1, out.data_ptr(), 3072, 0, 0, st))
def load_gemm0(st): L.check(lib.vc_gemm_bf16(A.data_ptr(), W0.data_ptr(), bias.data_ptr(), M, 2304, 768, 0, out.data_ptr(), 2304, 0, 0, st))
def load_ln(st): L.check(lib.vc_layernorm_f32_bf16(x.data_ptr(), g.data_ptr(), b0.data_ptr(), xn.data_ptr(), M, 768, 1e-6, st))
def load_att(st): L.check(lib.vc_vit_attention(qkv.data_ptr(), att.data_ptr(), 1024, 197, 12, 64, st))

for name, fn, reps in [("gemm gelu", load_gemm1, 60), ("gemm bias", load_gemm0, 80), ("layernorm", load_ln, 200), ("attention", load_att, 120)]:
    torch.cuda.synchronize()
    e0, e1, d0, d1 = ev(), ev(), ev(), ev()
    with torch.cuda.stream(E):
        e0.record()
        for _ in range(reps):
            fn(E.cuda_stream)
        e1.record()
    with torch.cuda.stream(D):
        d0.record()
        m.greedy_ids(prefix, None, n_new)
        d1.record()
    torch.cuda.synchronize()
    print(f"decode under a loop of {name:10s}: {d0.elapsed_time(d1):7.2f} ms (alone ~14.7)   load loop {e0.elapsed_time(e1):7.2f} ms")

# ---- graph replay vs eager launches of the same chain under a GEMM loop
for use_graph in (True, False):
    for _ in range(2):
        m.greedy_ids(prefix, None, n_new, use_graph=use_graph)
    torch.cuda.synchronize()
    e0, e1, d0, d1 = ev(), ev(), ev(), ev()
    with torch.cuda.stream(E):
        e0.record()
        for _ in range(60):
            load_gemm1(E.cuda_stream)
        e1.record()
    with torch.cuda.stream(D):
        d0.record()
        m.greedy_ids(prefix, None, n_new, use_graph=use_graph)
        d1.record()
    torch.cuda.synchronize()
    print(f"decode ({'graph' if use_graph else 'eager'}) under a loop of gemm gelu: {d0.elapsed_time(d1):7.2f} ms   load loop {e0.elapsed_time(e1):7.2f} ms")
