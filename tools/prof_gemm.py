"""Times the ViT GEMM shapes (target of `ncu --set full -k regex:gemm_tcgen05`).  `--sweep` also varies M so that
the A operand does / does not fit the 126 MB L2."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa: F401
from vcb200 import lib as L

lib = L.load()
st = torch.cuda.current_stream().cuda_stream
sweep = "--sweep" in sys.argv
Ms = [64 * 197, 128 * 197, 256 * 197] if sweep else [256 * 197]
iters = 20 if sweep else 5
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for M in Ms:
    for (N, K, mode) in [(3072, 768, 1), (768, 3072, 3), (2304, 768, 0), (768, 768, 3)]:
        A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
        W = (torch.randn(N, K, device="cuda") * 0.05).to(torch.bfloat16)
        bias = torch.randn(N, device="cuda")
        out = torch.zeros(M, N, device="cuda", dtype=torch.float32 if mode == 3 else torch.bfloat16)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(2):
            L.check(lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, mode, out.data_ptr(), N, 0, 0, st))
        e0.record()
        for _ in range(iters):
            L.check(lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, mode, out.data_ptr(), N, 0, 0, st))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        line = f"gemm M={M} N={N} K={K} mode={mode}: {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s"
        if sweep:   # cold: flush L2 before each launch, time each launch alone
            tot = 0.0
            for _ in range(5):
                flush.fill_(1)
                e0.record()
                L.check(lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, mode, out.data_ptr(), N, 0, 0, st))
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            line += f"   cold-L2 {2.0 * M * N * K / (tot / 5) / 1e9:.1f} TFLOP/s"
        print(line, flush=True)
