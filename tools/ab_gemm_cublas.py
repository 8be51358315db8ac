"""The encoder's GEMM shapes at cfg2's row count (1024 frames x 197 rows): this repo's tcgen05 kernel (bias epilogue, bf16 out)
against cuBLAS through torch (`F.linear`, bias in cuBLASLt's epilogue), alternated in one process and each timed over a sustained
loop, so that both run at the clock the power cap allows.  Says how much head-room the main loop has left on these shapes
(K = 768 is short: one output tile per 12 K blocks)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.nn.functional as F
import vcb200  # noqa: F401
from vcb200 import lib as L

lib = L.load()
st = torch.cuda.current_stream().cuda_stream
M = 1024 * 197
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 150
for (N, K, name) in [(2304, 768, "qkv"), (3072, 768, "fc1"), (768, 3072, "fc2"), (768, 768, "proj")]:
    A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    bias_b = bias.to(torch.bfloat16)
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)

    def ours():
        L.check(lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, 0, out.data_ptr(), N, 0, 0, st))

    def cublas():
        F.linear(A, W, bias_b)

    res = {"ours": [], "cublas": []}
    for rep in range(2):
        for nm, fn in (("ours", ours), ("cublas", cublas)):
            for _ in range(5):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[nm].append(2.0 * M * N * K / (e0.elapsed_time(e1) / iters) / 1e9)
    print(f"{name:5s} M={M} N={N} K={K}:  ours " + " ".join(f"{x:7.1f}" for x in res["ours"]) + "   cuBLAS " +
          " ".join(f"{x:7.1f}" for x in res["cublas"]) + "  TFLOP/s", flush=True)
