"""Runs the ViT GEMM shapes of one 256-frame chunk a few times (target of `ncu --set full -k regex:gemm_tcgen05`)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa: F401
from vcb200 import lib as L

lib = L.load()
M = 256 * 197
st = torch.cuda.current_stream().cuda_stream
for (N, K, mode) in [(3072, 768, 1), (768, 3072, 3), (2304, 768, 0)]:
    A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if mode == 3 else torch.bfloat16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(2):
        L.check(lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, mode, out.data_ptr(), N, 0, 0, st))
    e0.record()
    for _ in range(5):
        L.check(lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, mode, out.data_ptr(), N, 0, 0, st))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"gemm M={M} N={N} K={K} mode={mode}: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s")
