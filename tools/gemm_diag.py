"""Bring-up diagnostics for the tcgen05 GEMM (run on the GPU box).

python tools/gemm_diag.py            # default encoding on a ladder of shapes, then a sweep of
                                     # descriptor variants (each in its own process: a trap or a
                                     # hang in one variant must not poison the others)
python tools/gemm_diag.py one <desc_hi> <k_adv> <idesc> <M> <N> <K>
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def run_one(desc_hi: int, k_adv: int, idesc: int, M: int, N: int, K: int) -> None:
    import torch
    import vcb200  # noqa: F401
    from vcb200 import lib as L

    lib = L.load()
    lib.vc_debug_gemm_override(desc_hi, k_adv, idesc)
    torch.manual_seed(0)
    A = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * 0.5).to(torch.bfloat16)
    out = torch.full((M, N), 123.0, device="cuda", dtype=torch.float32)
    rc = lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), 0, M, N, K, 4, out.data_ptr(), N, 0, 0, torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        print("  rc", rc, lib.vc_last_error().decode())
        return
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    err = (out - ref).abs()
    bad = err > 0.05
    print(f"  M={M} N={N} K={K}: max_err={err.max().item():.4g} bad_frac={bad.float().mean().item():.4f} "
          f"untouched={(out == 123.0).float().mean().item():.4f}")
    if bad.any():
        rows = bad.any(1).nonzero().flatten().tolist()
        cols = bad.any(0).nonzero().flatten().tolist()
        print("   bad rows (first 24):", rows[:24], "count", len(rows))
        print("   bad cols (first 24):", cols[:24], "count", len(cols))
        # does each K=16 slice contribute correctly?  compare against partial sums
        for ks in range(0, min(K, 64), 16):
            part = A[:, ks:ks + 16].float() @ W[:, ks:ks + 16].float().t()
            print(f"   corr with K-slice {ks // 16}: {torch.corrcoef(torch.stack([out.flatten(), part.flatten()]))[0, 1].item():.3f}")
        print("   out[0,:4]", out[0, :4].tolist(), "ref[0,:4]", ref[0, :4].tolist())


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        run_one(*[int(v, 0) for v in sys.argv[2:8]])
        return
    base_hi = (1 << 16) | ((1024 >> 4) << 32) | (1 << 46) | (2 << 61)
    variants = [
        ("default sw128 v1 lbo1 sbo1024 kadv2", base_hi, 2, 0),
        ("version bit off", base_hi & ~(1 << 46), 2, 0),
        ("lbo 0", base_hi & ~(1 << 16), 2, 0),
        ("sbo 512", (base_hi & ~(0x3FFF << 32)) | ((512 >> 4) << 32), 2, 0),
        ("kadv 4", base_hi, 4, 0),
    ]
    shapes = [(128, 256, 64), (128, 256, 256), (300, 512, 768), (64, 768, 768)]
    for name, hi, kadv, idesc in variants:
        print(f"== {name}", flush=True)
        for (M, N, K) in (shapes if name.startswith("default") else shapes[:1]):
            try:
                r = subprocess.run([sys.executable, __file__, "one", hex(hi), str(kadv), str(idesc), str(M), str(N), str(K)],
                                   capture_output=True, text=True, timeout=180)
                print(r.stdout.rstrip() or "  (no output)")
                if r.returncode != 0:
                    print("  exit", r.returncode, r.stderr.strip().splitlines()[-3:])
            except subprocess.TimeoutExpired:
                print("  TIMEOUT (hang)")
        sys.stdout.flush()


if __name__ == "__main__":
    main()
