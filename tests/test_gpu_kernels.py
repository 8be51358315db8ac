"""Per-kernel parity on the B200 through the C ABI (ctypes), against fp32 torch references
and the CPU oracle.  Integer / byte / index work must be bit-exact; floating point states
its tolerance next to the assert."""
import ctypes as C

import pytest
import torch

import vcb200  # noqa: F401
from vcb200 import lib as L
from oracle import vc_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


@pytest.fixture(scope="module")
def lib():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return L.load()


def _stream():
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------ preprocessing (bit-exact)
@pytest.mark.parametrize("n,size", [(1, 224), (5, 224), (0, 224)])
def test_preprocess_chw_bit_exact(lib, n, size):
    g = torch.Generator().manual_seed(7)
    frames = torch.randint(0, 256, (n, size, size, 3), generator=g, dtype=torch.uint8)
    ref = O.preprocess_u8(frames).to(torch.bfloat16)          # reference tensor, bf16-rounded
    lut = O.normalize_lut().to(DEV)
    out = torch.empty(n, 3, size, size, device=DEV, dtype=torch.bfloat16)
    L.check(lib.vc_preprocess_u8(frames.to(DEV).data_ptr(), lut.data_ptr(), out.data_ptr(), n, size, size, 0, 16, 768, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), ref)


@pytest.mark.parametrize("patch,k_pad", [(16, 768), (14, 640)])
def test_preprocess_patch_major_is_a_permutation_of_chw(lib, patch, k_pad):
    n, size = 3, 224
    g = torch.Generator().manual_seed(11)
    frames = torch.randint(0, 256, (n, size, size, 3), generator=g, dtype=torch.uint8)
    ref = O.preprocess_u8(frames).to(torch.bfloat16)          # [n,3,H,W]
    gsz = size // patch
    # im2col of the reference tensor: rows (frame, py, px), columns (c, i, j) — Conv2d weight order
    cols = ref.view(n, 3, gsz, patch, gsz, patch).permute(0, 2, 4, 1, 3, 5).reshape(n * gsz * gsz, 3 * patch * patch)
    lut = O.normalize_lut().to(DEV)
    out = torch.full((n * gsz * gsz, k_pad), 7.0, device=DEV, dtype=torch.bfloat16)
    L.check(lib.vc_preprocess_u8(frames.to(DEV).data_ptr(), lut.data_ptr(), out.data_ptr(), n, size, size, 1, patch, k_pad, _stream()))
    torch.cuda.synchronize()
    out = out.cpu()
    assert torch.equal(out[:, : 3 * patch * patch], cols)
    assert bool((out[:, 3 * patch * patch:] == 0).all())      # K padding is zero


def test_patchify_f32_matches_preprocess(lib):
    n, size = 2, 224
    frames = torch.randint(0, 256, (n, size, size, 3), dtype=torch.uint8)
    video = O.preprocess_u8(frames)
    a = torch.empty(n * 196, 768, device=DEV, dtype=torch.bfloat16)
    b = torch.empty_like(a)
    lut = O.normalize_lut().to(DEV)
    L.check(lib.vc_preprocess_u8(frames.to(DEV).data_ptr(), lut.data_ptr(), a.data_ptr(), n, size, size, 1, 16, 768, _stream()))
    L.check(lib.vc_patchify_f32(video.to(DEV).data_ptr(), b.data_ptr(), n, size, size, 16, 768, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(a, b)


# ------------------------------------------------------------------ tcgen05 GEMM
def _gemm_ref(A, W, bias):
    return A.float() @ W.float().t() + (bias if bias is not None else 0.0)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 768), (197, 768, 768), (1000, 2304, 768), (64, 768, 3072),
                                   (333, 3072, 768), (5, 256, 128), (4096, 768, 3072)])
def test_gemm_bias_bf16(lib, M, N, K):
    torch.manual_seed(M + N + K)
    A = (torch.randn(M, K, device=DEV) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    out = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    L.check(lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, 0, out.data_ptr(), N, 0, 0, _stream()))
    torch.cuda.synchronize()
    ref = _gemm_ref(A, W, bias)
    err = (out.float() - ref).abs().max().item()
    # fp32 accumulate of bf16 products, output rounded to bf16: |ref| <~ 8 -> half-ulp 2^-6 plus accumulation noise
    assert err < 0.05, err


@pytest.mark.parametrize("mode", [1, 2])
def test_gemm_gelu_epilogues(lib, mode):
    M, N, K = 300, 512, 256
    torch.manual_seed(mode)
    A = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV) * 0.1).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV) * 0.1
    out = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    L.check(lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, mode, out.data_ptr(), N, 0, 0, _stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.gelu(_gemm_ref(A, W, bias), approximate="tanh" if mode == 2 else "none")
    assert (out.float() - ref).abs().max().item() < 0.03      # bf16 output rounding of |y| < 8


def test_gemm_residual_and_f32_epilogues(lib):
    M, N, K = 257, 768, 3072
    torch.manual_seed(3)
    A = (torch.randn(M, K, device=DEV) * 0.3).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV) * 0.02).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV)
    res = torch.randn(M, N, device=DEV)
    ref = res + _gemm_ref(A, W, bias)
    L.check(lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, 3, res.data_ptr(), N, 0, 0, _stream()))
    out32 = torch.zeros(M, N, device=DEV)
    L.check(lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), 0, M, N, K, 4, out32.data_ptr(), N, 0, 0, _stream()))
    torch.cuda.synchronize()
    assert (res - ref).abs().max().item() < 2e-3               # fp32 out: only accumulation-order noise
    assert (out32 - _gemm_ref(A, W, None)).abs().max().item() < 2e-3


@pytest.mark.parametrize("M,K", [(64, 768), (256, 768), (256, 3072), (320, 3072), (130, 768)])
def test_gemm_resid_stats_split_k(lib, M, K):
    """x += bf16(A W^T + b), xb = bf16(x), row statistics — with K walked by one CTA pair per tile and split over 2 / 4 pairs of a
    cluster (the few-row products of the GPT-2 chain): every variant against torch, the split ones against the unsplit one to
    accumulation-order noise, and the fused "last CTA" (mean, rstd) against torch's LayerNorm statistics."""
    N = 768
    torch.manual_seed(M + K)
    A = (torch.randn(M, K, device=DEV) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV) * 0.03).to(torch.bfloat16)
    bias = torch.randn(N, device=DEV) * 0.1
    x0 = torch.randn(M, N, device=DEV)
    delta = (A.float() @ W.float().t() + bias).to(torch.bfloat16).float()
    ref = x0 + delta
    outs = {}
    for ks in (1, 2, 4, 0):
        x = x0.clone()
        xb = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
        pst = torch.full((N // 32, M, 2), float("nan"), device=DEV)
        st = torch.zeros(M, 2, device=DEV)
        done = torch.zeros(1, device=DEV, dtype=torch.int32)
        L.check(lib.vc_gemm_resid_stats(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, N, K, x.data_ptr(), xb.data_ptr(), pst.data_ptr(),
                                        st.data_ptr(), done.data_ptr(), 1e-5, ks, _stream()))
        torch.cuda.synchronize()
        assert int(done.item()) == 0                              # the last CTA resets the counter
        # one bf16 ulp of the delta (|delta| <~ 4 -> 2^-6) where the accumulation order moved a rounding boundary
        assert (x - ref).abs().max().item() <= 0.04, ks
        assert torch.equal(xb, x.to(torch.bfloat16)), ks
        chunks = x.view(M, N // 32, 32)
        assert torch.allclose(pst[..., 0].t(), chunks.sum(-1), atol=2e-4, rtol=1e-5), ks
        assert torch.allclose(pst[..., 1].t(), (chunks * chunks).sum(-1), atol=2e-3, rtol=1e-5), ks
        mean = x.double().mean(-1)
        rstd = 1.0 / torch.sqrt(x.double().var(-1, unbiased=False) + 1e-5)
        assert torch.allclose(st[:, 0].double(), mean, atol=1e-5, rtol=1e-5), ks
        assert torch.allclose(st[:, 1].double(), rstd, atol=1e-5, rtol=1e-5), ks
        outs[ks] = x
    for ks in (2, 4):
        d = (outs[ks] - outs[1]).abs()
        assert d.max().item() <= 0.04 and (d > 0).float().mean().item() < 0.02, ks     # same values up to rare rounding flips


def test_gemm_patch_embed_epilogue(lib):
    frames, P, D, K = 3, 196, 768, 768
    M = frames * P
    torch.manual_seed(5)
    A = torch.randn(M, K, device=DEV).to(torch.bfloat16)
    W = (torch.randn(D, K, device=DEV) * 0.03).to(torch.bfloat16)
    bias = torch.randn(D, device=DEV)
    pos = torch.randn(P + 1, D, device=DEV)
    x = torch.full((frames * (P + 1), D), -5.0, device=DEV)
    L.check(lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), M, D, K, 5, x.data_ptr(), D, pos.data_ptr(), P, _stream()))
    torch.cuda.synchronize()
    ref = (_gemm_ref(A, W, bias).view(frames, P, D) + pos[1:].unsqueeze(0))
    got = x.view(frames, P + 1, D)
    assert (got[:, 1:] - ref).abs().max().item() < 2e-3
    assert bool((got[:, 0] == -5.0).all())                     # class-token rows untouched


def test_gemm_rejects_bad_k(lib):
    A = torch.zeros(8, 40, device=DEV, dtype=torch.bfloat16)
    W = torch.zeros(32, 40, device=DEV, dtype=torch.bfloat16)
    out = torch.zeros(8, 32, device=DEV, dtype=torch.bfloat16)
    rc = lib.vc_gemm_bf16(A.data_ptr(), W.data_ptr(), 0, 8, 32, 40, 0, out.data_ptr(), 32, 0, 0, _stream())
    assert rc != 0 and b"multiple of 64" in lib.vc_last_error()


# ------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("rows,dim,eps", [(1, 768, 1e-6), (197 * 3, 768, 1e-6), (64, 1024, 1e-5)])
def test_layernorm(lib, rows, dim, eps):
    torch.manual_seed(rows)
    x = torch.randn(rows, dim, device=DEV) * 2 + 0.3
    g = torch.randn(dim, device=DEV) * 0.1 + 1
    b = torch.randn(dim, device=DEV) * 0.05
    out = torch.empty(rows, dim, device=DEV, dtype=torch.bfloat16)
    L.check(lib.vc_layernorm_f32_bf16(x.data_ptr(), g.data_ptr(), b.data_ptr(), out.data_ptr(), rows, dim, eps, _stream()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.layer_norm(x, (dim,), g, b, eps)
    assert (out.float() - ref).abs().max().item() < 0.04       # bf16 rounding of |y| < 8


# ------------------------------------------------------------------ ViT attention
@pytest.mark.parametrize("frames,tokens,heads", [(1, 197, 12), (3, 197, 12), (2, 257, 16), (2, 50, 2), (1, 64, 1)])
def test_vit_attention(lib, frames, tokens, heads):
    D = heads * 64
    torch.manual_seed(tokens)
    qkv = torch.randn(frames * tokens, 3 * D, device=DEV).to(torch.bfloat16)
    out = torch.zeros(frames * tokens, D, device=DEV, dtype=torch.bfloat16)
    L.check(lib.vc_vit_attention(qkv.data_ptr(), out.data_ptr(), frames, tokens, heads, 64, _stream()))
    torch.cuda.synchronize()
    q, k, v = qkv.float().view(frames, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1) @ v
    ref = ref.transpose(1, 2).reshape(frames * tokens, D)
    assert (out.float() - ref).abs().max().item() < 0.03       # P rounded to bf16 before PV + bf16 output


def test_vit_attention_tcgen05_agrees_with_mma_sync_kernel(lib):
    """The tcgen05 kernel (S/P/O in TMEM) and the mma.sync kernel it replaced for tokens <= 256 compute the same
    softmax(QK^T/8)V: both round P to bf16 before PV, so they agree to bf16 output rounding."""
    frames, tokens, heads = 5, 197, 12
    D = heads * 64
    torch.manual_seed(3)
    qkv = (torch.randn(frames * tokens, 3 * D, device=DEV) * 1.5).to(torch.bfloat16)
    a = torch.zeros(frames * tokens, D, device=DEV, dtype=torch.bfloat16)
    b = torch.zeros_like(a)
    L.check(lib.vc_vit_attention(qkv.data_ptr(), a.data_ptr(), frames, tokens, heads, 64, _stream()))
    L.check(lib.vc_vit_attention_mma_sync(qkv.data_ptr(), b.data_ptr(), frames, tokens, heads, 64, _stream()))
    torch.cuda.synchronize()
    assert (a.float() - b.float()).abs().max().item() < 0.02
    # frames are independent: the result of frame 2 does not depend on what surrounds it (the kernel's K/V tiles run
    # over the frame boundary and rely on masking)
    c = torch.zeros(tokens, D, device=DEV, dtype=torch.bfloat16)
    one = qkv[2 * tokens:3 * tokens].contiguous()
    L.check(lib.vc_vit_attention(one.data_ptr(), c.data_ptr(), 1, tokens, heads, 64, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(c, a[2 * tokens:3 * tokens])


# ------------------------------------------------------------------ pool / prefix and the two CuPy-hook operators
def test_pool_prefix_matches_oracle(lib):
    B, T, D, Vd, Pn, H = 3, 4, 768, 256, 4, 768
    torch.manual_seed(1)
    cls = torch.randn(B * T, D)
    sd = {"encoder.proj.weight": torch.randn(Vd, D) * 0.02, "encoder.proj.bias": torch.randn(Vd) * 0.02,
          "decoder.mapper.0.weight": torch.randn(Pn * H, Vd) * 0.03, "decoder.mapper.0.bias": torch.randn(Pn * H) * 0.02}
    feat_ref = torch.nn.functional.linear(cls.view(B, T, D).mean(1), sd["encoder.proj.weight"], sd["encoder.proj.bias"])
    prefix_ref = O.visual_prefix(sd, feat_ref, 0.6, 0.4, Pn)
    d = {k: v.to(DEV) for k, v in sd.items()}
    feat = torch.empty(B, Vd, device=DEV)
    prefix = torch.empty(B, Pn * H, device=DEV)
    L.check(lib.vc_pool_prefix(cls.to(DEV).data_ptr(), B, T, D, d["encoder.proj.weight"].data_ptr(), d["encoder.proj.bias"].data_ptr(),
                               Vd, 0.6, 0.4, d["decoder.mapper.0.weight"].data_ptr(), d["decoder.mapper.0.bias"].data_ptr(), Pn * H,
                               feat.data_ptr(), prefix.data_ptr(), _stream()))
    torch.cuda.synchronize()
    assert (feat.cpu() - feat_ref).abs().max().item() < 1e-5   # fp32 throughout, summation order only
    assert (prefix.cpu().view(B, Pn, H) - prefix_ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("gap", [0, 1])
def test_vit_pool_temporal_hook(lib, gap):
    """Same contract as core/operators/cupy_vit_pool.py:127-186 (cls / gap)."""
    bsz, T, N, Cc = 2, 3, 17, 96
    feat = torch.randn(bsz * T, N, Cc, device=DEV)
    out = torch.empty(bsz, Cc, device=DEV)
    L.check(lib.vc_vit_pool_temporal(feat.data_ptr(), 0, bsz, T, N, Cc, gap, out.data_ptr(), _stream()))
    torch.cuda.synchronize()
    f = feat.view(bsz, T, N, Cc)
    ref = f[:, :, 1:, :].mean(dim=(1, 2)) if gap else f[:, :, 0, :].mean(dim=1)
    assert (out - ref).abs().max().item() < 1e-5


def test_linear_bias_hook(lib):
    """Same contract as cupy_linear_mapper.py `linear_bias_f32` (:14-40)."""
    x = torch.randn(5, 256, device=DEV)
    w = torch.randn(3072, 256, device=DEV) * 0.05
    b = torch.randn(3072, device=DEV)
    y = torch.empty(5, 3072, device=DEV)
    L.check(lib.vc_linear_bias_f32(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), 5, 256, 3072, _stream()))
    torch.cuda.synchronize()
    assert (y - torch.nn.functional.linear(x, w, b)).abs().max().item() < 1e-4


# ------------------------------------------------------------------ token selection (bit-exact)
def test_argmax_bit_exact_with_ties(lib):
    torch.manual_seed(0)
    rows, V = 37, 50257
    logits = torch.randn(rows, V)
    logits[3, 100] = logits[3, 40000] = 9.0                   # tie -> lowest index
    logits[5, :] = -1.5                                        # all equal -> index 0
    logits[7, V - 1] = 50.0
    out = torch.empty(rows, device=DEV, dtype=torch.int32)
    L.check(lib.vc_argmax_f32(logits.to(DEV).data_ptr(), rows, V, out.data_ptr(), _stream()))
    torch.cuda.synchronize()
    assert out.cpu().tolist() == torch.argmax(logits, dim=-1).tolist()
    assert out[3].item() == 100 and out[5].item() == 0 and out[7].item() == V - 1


# ------------------------------------------------------------------ decode-step skinny GEMM (split-K partials)
@pytest.mark.parametrize("M,N,K,ks", [(64, 2304, 768, 0), (64, 768, 3072, 0), (3, 768, 768, 12), (128, 3072, 768, 3), (64, 50432, 768, 1),
                                      (100, 1024, 4096, 0)])
def test_skinny_gemm_partials_sum_to_product(lib, M, N, K, ks):
    torch.manual_seed(M + N)
    x = (torch.randn(M, K, device=DEV) * 0.5).to(torch.bfloat16)
    W = (torch.randn(N, K, device=DEV) * 0.05).to(torch.bfloat16)
    ksplit = ks or lib.vc_skinny_ksplit(N, K)
    P = torch.full((ksplit, M, N), 7.0, device=DEV)
    L.check(lib.vc_skinny_gemm_partial(x.data_ptr(), W.data_ptr(), P.data_ptr(), M, N, K, ksplit, _stream()))
    torch.cuda.synchronize()
    ref = x.float() @ W.float().t()
    assert (P.sum(0) - ref).abs().max().item() < 2e-3          # fp32 accumulate; order differs from cuBLAS only
    P2 = torch.empty_like(P)
    L.check(lib.vc_skinny_gemm_partial(x.data_ptr(), W.data_ptr(), P2.data_ptr(), M, N, K, ksplit, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(P, P2)                                  # deterministic: no atomics


# ------------------------------------------------------------------ HF logits processors + top-K continuations (bit-exact ids)
@pytest.mark.parametrize("raw", [0, 1])
def test_beam_step_matches_oracle_processors(lib, raw):
    torch.manual_seed(9)
    B, nb, V, max_len, cur_len = 3, 4, 50257, 12, 7
    K = 1 if raw else 2 * nb
    rows = B * (1 if raw else nb)
    per_item = 1 if raw else nb
    logits = torch.randn(rows, V) * 2
    seqs = torch.randint(0, V, (rows, max_len))
    seqs[:, 3] = seqs[:, 1]; seqs[:, 4] = seqs[:, 2]; seqs[:, 5] = seqs[:, 1]; seqs[:, 6] = seqs[:, 2]   # repeated bigram -> banned token
    running = torch.randn(rows) * 3
    if raw:
        sc = O.apply_processors(logits, seqs[:, :cur_len], cur_len, repetition_penalty=1.1, no_repeat_ngram_size=3, min_new_tokens=8, eos=50256)
        ref_idx = sc.argmax(-1).view(rows, 1)
        ref_val = sc.max(-1).values.view(rows, 1)
    else:
        lp = torch.log_softmax(logits, -1)
        sc = O.apply_processors(lp, seqs[:, :cur_len], cur_len, repetition_penalty=1.1, no_repeat_ngram_size=3, min_new_tokens=8, eos=50256)
        cand = (sc.view(B, nb, V) + running.view(B, nb, 1)).view(B, nb * V)
        ref_val, ref_idx = torch.topk(cand, K, dim=1)
    d = lambda t, dt: t.to(DEV).to(dt).contiguous()
    cs = torch.empty(rows, K, device=DEV); ct = torch.empty(rows, K, device=DEV, dtype=torch.int32)
    ts = torch.empty(rows // per_item, K, device=DEV); ti = torch.empty(rows // per_item, K, device=DEV, dtype=torch.int32)
    lg, sq, rn = d(logits, torch.float32), d(seqs, torch.int32), d(running, torch.float32)
    L.check(lib.vc_beam_step(lg.data_ptr(), V, V, rows, per_item, sq.data_ptr(), max_len, cur_len, rn.data_ptr(), 1.1, 3, 8, 50256, raw, K,
                             cs.data_ptr(), ct.data_ptr(), ts.data_ptr(), ti.data_ptr(), _stream()))
    torch.cuda.synchronize()
    assert ti.cpu().tolist() == ref_idx.tolist()
    assert (ts.cpu() - ref_val).abs().max().item() < 2e-5


def test_beam_reorder_is_index_select(lib):
    n_seq, s_max, upto = 12, 9, 6
    slot_in = torch.randint(0, n_seq, (n_seq, s_max), dtype=torch.int32)
    src = torch.randint(0, n_seq, (n_seq,), dtype=torch.int32)
    out = torch.full((n_seq, s_max), -1, dtype=torch.int32).to(DEV)
    L.check(lib.vc_beam_reorder(slot_in.to(DEV).data_ptr(), out.data_ptr(), src.to(DEV).data_ptr(), n_seq, s_max, upto, _stream()))
    torch.cuda.synchronize()
    ref = torch.full((n_seq, s_max), -1, dtype=torch.int32)
    ref[:, :upto] = slot_in[src.long(), :upto]
    assert torch.equal(out.cpu(), ref)


def test_resize_byte_exact_vs_pillow_golden_and_oracle(lib, golden_dir):
    """vc_resize_bilinear_u8 against the reference's PIL resize (golden fixtures) and against the oracle on fresh frames:
    batched, odd sizes, row counts that are not multiples of the CTA's row group, upscaling — byte for byte."""
    import numpy as np
    from vcb200.resample import FrameResizer
    rs = FrameResizer(DEV, 224, 224)
    z = np.load(golden_dir / "resize.npz")
    for k in [k for k in z.files if k.startswith("src_")]:
        got = rs(torch.from_numpy(z[k]).to(DEV))
        assert torch.equal(got.cpu(), torch.from_numpy(z["dst_" + k[4:]])), k
    g = torch.Generator().manual_seed(3)
    for shape in [(3, 2, 97, 131, 3), (1, 360, 480, 3), (5, 224, 321, 3), (2, 333, 224, 3), (2, 50, 40, 3)]:
        x = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
        got = rs(x.to(DEV))
        assert got.shape == shape[:-3] + (224, 224, 3)
        assert torch.equal(got.cpu(), O.resize_bilinear_u8(x, 224, 224)), shape
    x = torch.randint(0, 256, (2, 224, 224, 3), generator=g, dtype=torch.uint8)
    assert torch.equal(rs(x.to(DEV)).cpu(), x)


def test_decode_chain_fused_fc_variant_matches(lib):
    """skinny_gemm_gelu (fc1 + bias + gelu_new in one kernel, opt-in via VC_DECODE_FUSED_FC) against the default pair of
    kernels through the public forward: same logits to fp32 summation-order noise."""
    import os
    from vcb200 import synthetic
    from vcb200.model import B200CaptionModel
    a = synthetic.ARCHS["tiny"]
    m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), DEV, vit_heads=a.vit_heads, gpt_heads=a.gpt_heads)
    g = torch.Generator().manual_seed(4)
    prefix = (torch.randn(70, a.prefix_len, a.gpt_dim, generator=g) * 0.3).to(DEV)
    forced = torch.randint(0, 50000, (70, 4), generator=g).int().to(DEV)
    outs = []
    for flag in (None, "1"):
        if flag:
            os.environ["VC_DECODE_FUSED_FC"] = flag
        else:
            os.environ.pop("VC_DECODE_FUSED_FC", None)
        try:
            ids, lens, lg = m.greedy_ids(prefix, None, 4, forced_ids=forced, keep_logits=True, use_graph=False)
            torch.cuda.synchronize()
            outs.append((ids.cpu().clone(), lg.float().cpu().clone()))
        finally:
            os.environ.pop("VC_DECODE_FUSED_FC", None)
    assert (outs[0][1] - outs[1][1]).abs().max().item() < 2e-2
    assert (outs[0][0] == outs[1][0]).float().mean().item() >= 0.97      # only a near-tie may flip an argmax (different fp32 summation order)


def test_lm_head_tcgen05_matches_mma_sync_version(lib):
    """dc_lmhead_tc_kernel (tcgen05, the default) against dc_lmhead_kernel (mma.sync, VC_LMHEAD_TC=0) through the public greedy
    call with teacher forcing: same logits to fp32 summation-order noise, same candidates -> same ids unless a near-tie flips;
    61 sequences (rows 61-63 of the A box are out of bounds -> zero filled), a full 64, and 100 sequences on the same chain
    (VC_DECODE_CHAIN=2: two row tiles, blockIdx.y = 1 reads rows 64-99)."""
    import os
    from vcb200 import synthetic
    from vcb200.model import B200CaptionModel
    a = synthetic.ARCHS["tiny"]
    m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), DEV, vit_heads=a.vit_heads, gpt_heads=a.gpt_heads)
    g = torch.Generator().manual_seed(11)
    try:
        for n_seq in (61, 64, 100):
            if n_seq > 64:
                os.environ["VC_DECODE_CHAIN"] = "2"          # read per call: the few-row chain beyond 64 rows
            prefix = (torch.randn(n_seq, a.prefix_len, a.gpt_dim, generator=g) * 0.3).to(DEV)
            forced = torch.randint(0, 50000, (n_seq, 3), generator=g).int().to(DEV)
            outs = []
            for flag in (None, "0"):
                if flag:
                    os.environ["VC_LMHEAD_TC"] = flag
                else:
                    os.environ.pop("VC_LMHEAD_TC", None)
                try:
                    ids, lens, lg = m.greedy_ids(prefix, None, 3, forced_ids=forced, keep_logits=True, use_graph=False)
                    torch.cuda.synchronize()
                    outs.append((ids.cpu().clone(), lg.float().cpu().clone(), lens.cpu().clone()))
                finally:
                    os.environ.pop("VC_LMHEAD_TC", None)
            # steps 1.. run on the few-row chain (the 5-position prefill of > 64 rows takes the GEMM chain in both runs)
            assert (outs[0][1] - outs[1][1]).abs().max().item() < 2e-3, n_seq
            assert (outs[0][0] == outs[1][0]).float().mean().item() >= 0.98, n_seq
            # ids are the argmax of the logits each version produced itself (ties -> lowest index); a row that has emitted eos keeps
            # eos from then on, whatever its logits say (benchmark_baseline.py:212-224)
            for ids, lg, lens in outs:
                live = torch.arange(3)[None, :] < lens[:, None].long()
                am = lg.argmax(-1).t()
                assert torch.equal(ids.long()[live], am[live]), n_seq
                assert live[:, 1:].float().mean().item() > 0.9
    finally:
        os.environ.pop("VC_DECODE_CHAIN", None)


def test_beam_update_kernel_equals_torch_bookkeeping():
    """vc_beam_update (one kernel per step) against the torch formulation of transformers' `_beam_search` state update
    (_get_running_beams_for_next_iteration, _update_finished_beams, _check_early_stop_heuristic; the oracle's beam_search uses the
    same lines), step by step on random candidate lists with eos hits, for 7 videos x 4 beams."""
    import ctypes as C
    from vcb200 import lib as L
    lib = L.load()
    B, nb, mx, V, eos = 7, 4, 12, 1000, 999
    K = 2 * nb
    NEG = -1.0e9
    g = torch.Generator().manual_seed(0)
    i32 = lambda *s: torch.zeros(*s, device=DEV, dtype=torch.int32)
    f32 = lambda *s: torch.zeros(*s, device=DEV, dtype=torch.float32)
    bufs = dict(running_scores=f32(B, nb), running_seqs=i32(B * nb, mx), fin_seqs=i32(B, nb, mx), fin_scores=f32(B, nb), fin_done=i32(B, nb),
                fin_len=i32(B, nb), unsatisfied=i32(B), flags=i32(mx + 1, 2), stopped=i32(1), src_rows=i32(B * nb), next_tok=i32(B * nb))
    bs = L.VcBeamState()
    bs.B, bs.nb, bs.max_len, bs.eos = B, nb, mx, eos
    for k, v in bufs.items():
        setattr(bs, k, v.data_ptr())
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.vc_beam_init(C.byref(bs), st))
    # torch reference state (CPU)
    running_scores = torch.zeros(B, nb); running_scores[:, 1:] = NEG
    running_seqs = torch.full((B, nb, mx), eos, dtype=torch.int64)
    fin_seqs = running_seqs.clone(); fin_scores = torch.full((B, nb), NEG)
    fin_done = torch.zeros(B, nb, dtype=torch.bool); fin_len = torch.zeros(B, nb, dtype=torch.int64)
    unsat = torch.ones(B, 1, dtype=torch.bool); stopped = False
    in_top = torch.arange(K).view(1, K) < nb
    for cur_len in range(mx):
        top_scores = torch.sort(torch.randn(B, K, generator=g) * 2 - 3 * (cur_len + 1), dim=1, descending=True).values
        # candidates continue live beams only (a -1e9 filler beam can never reach the top 2*nb in a real search, and the order of
        # fillers among themselves is not defined)
        n_live = (running_scores > NEG / 2).sum(dim=1, keepdim=True).clamp(min=1)
        top_beam = (torch.rand(B, K, generator=g) * n_live).floor().long().clamp(max=nb - 1)
        top_tok = torch.randint(0, V - 1, (B, K), generator=g)
        top_tok[torch.rand(B, K, generator=g) < (0.25 if cur_len >= 2 else 0.0)] = eos
        top_idx = (top_beam * V + top_tok).int()
        d_scores, d_idx = top_scores.to(DEV), top_idx.to(DEV)           # keep the device copies alive across the launch
        L.check(lib.vc_beam_update(C.byref(bs), d_scores.data_ptr(), d_idx.data_ptr(), V, cur_len, 1.0, st))
        torch.cuda.synchronize()
        # ---- torch formulation
        cand_seqs = torch.gather(running_seqs, 1, top_beam.unsqueeze(-1).expand(-1, -1, mx)).clone()
        cand_seqs[:, :, cur_len] = top_tok
        new_len = cur_len + 1
        hit = (top_tok == eos) | (new_len >= mx)
        run_rank = top_scores + hit.float() * NEG
        running_scores, pick = torch.topk(run_rank, nb, dim=1)
        running_seqs = torch.gather(cand_seqs, 1, pick.unsqueeze(-1).expand(-1, -1, mx))
        running_beam, running_tok = torch.gather(top_beam, 1, pick), torch.gather(top_tok, 1, pick)
        newly = hit & in_top
        f = top_scores / (float(new_len) ** 1.0)
        f = f + (~unsat).float() * NEG
        f = f + (~newly).float() * NEG
        ms, mseq = torch.cat([fin_scores, f], 1), torch.cat([fin_seqs, cand_seqs], 1)
        md, ml = torch.cat([fin_done, newly], 1), torch.cat([fin_len, torch.full((B, K), new_len)], 1)
        n_scores, sel = torch.topk(ms, nb, dim=1)
        if not stopped:
            fin_scores, fin_seqs = n_scores, torch.gather(mseq, 1, sel.unsqueeze(-1).expand(-1, -1, mx))
            fin_done, fin_len = torch.gather(md, 1, sel), torch.gather(ml, 1, sel)
        best_running = running_scores[:, :1] / (float(new_len) ** 1.0)
        worst = torch.where(fin_done, fin_scores.min(dim=1, keepdim=True).values, torch.full_like(fin_scores, NEG))
        unsat = unsat & (best_running > worst).any(dim=-1, keepdim=True)
        stop_next = stopped or not (bool(unsat.any()) and not bool(hit.all()))
        # ---- compare (valid entries only: ties among -1e9 fillers have no defined order)
        ok_run = running_scores > NEG / 2
        assert torch.equal(bufs["running_scores"].cpu()[ok_run], running_scores[ok_run]), cur_len
        got_seqs = bufs["running_seqs"].cpu().view(B, nb, mx).long()
        assert torch.equal(got_seqs[ok_run], running_seqs[ok_run]), cur_len
        assert torch.equal(bufs["src_rows"].cpu().view(B, nb)[ok_run].long(), (running_beam + torch.arange(B).view(B, 1) * nb)[ok_run]), cur_len
        assert torch.equal(bufs["next_tok"].cpu().view(B, nb)[ok_run].long(), running_tok[ok_run]), cur_len
        ok_fin = fin_scores > NEG / 2
        assert torch.equal(bufs["fin_scores"].cpu()[ok_fin], fin_scores[ok_fin]), cur_len
        assert torch.equal(bufs["fin_seqs"].cpu().long()[ok_fin], fin_seqs[ok_fin]), cur_len
        assert torch.equal(bufs["fin_done"].cpu().bool()[ok_fin], fin_done[ok_fin]) and torch.equal(bufs["fin_len"].cpu().long()[ok_fin], fin_len[ok_fin]), cur_len
        assert torch.equal(bufs["unsatisfied"].cpu().bool().view(B, 1), unsat), cur_len
        stopped = stop_next
    ids = i32(B, mx); lens = i32(B)
    L.check(lib.vc_beam_finalize(C.byref(bs), ids.data_ptr(), lens.data_ptr(), st))
    torch.cuda.synchronize()
    assert lens.cpu().long().tolist() == fin_len[:, 0].tolist()
    for b in range(B):
        n = int(fin_len[b, 0])
        assert ids[b, :n].cpu().long().tolist() == fin_seqs[b, 0, :n].tolist() and ids[b, n:].eq(eos).all()
