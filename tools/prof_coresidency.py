"""Can another kernel's CTAs run beside the resident CTAs of the tcgen05 GEMM?  A chain of 200 tiny kernels (one small
CTA each) on a high-priority stream, alone and under a loop of GEMM launches; the same with 148 x 128-thread CTAs."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import lib as L
lib = L.load()
M = 1024 * 197
A = (torch.randn(M, 768, device="cuda") * 0.5).to(torch.bfloat16)
W1 = (torch.randn(3072, 768, device="cuda") * 0.05).to(torch.bfloat16)
bias = torch.randn(3072, device="cuda")
out = torch.zeros(M, 3072, device="cuda", dtype=torch.bfloat16)
E, D = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)
tiny = torch.zeros(32, device="cuda")
mid = torch.zeros(148 * 128 * 4, device="cuda")


def gemm(st):
    L.check(lib.vc_gemm_bf16(A.data_ptr(), W1.data_ptr(), bias.data_ptr(), M, 3072, 768, 1, out.data_ptr(), 3072, 0, 0, st))


def ev():
    return torch.cuda.Event(enable_timing=True)


for t, name in [(tiny, "1 CTA x 32 thr"), (mid, "148 CTAs x 128 thr")]:
    for load in (False, True):
        torch.cuda.synchronize()
        e0, e1, d0, d1 = ev(), ev(), ev(), ev()
        if load:
            with torch.cuda.stream(E):
                e0.record()
                for _ in range(40):
                    gemm(E.cuda_stream)
                e1.record()
        with torch.cuda.stream(D):
            d0.record()
            for _ in range(200):
                t.add_(1.0)
            d1.record()
        torch.cuda.synchronize()
        print(f"chain of 200 kernels ({name}) {'under GEMM loop' if load else 'alone          '}: {d0.elapsed_time(d1):8.2f} ms"
              + (f"   (GEMM loop {e0.elapsed_time(e1):.1f} ms)" if load else ""))
