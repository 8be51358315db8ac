"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/vcb200.h
declares; the product package never touches oracle/; host-side packing logic."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

import vcb200  # noqa: F401
from vcb200 import lib as L
from vcb200 import synthetic

ROOT = Path(__file__).resolve().parents[1]


def test_header_symbols_are_exported():
    header = (ROOT / "include" / "vcb200.h").read_text()
    declared = set(re.findall(r"\b(vc_[a-z0-9_]+)\s*\(", header)) - {"vc_stream_t"}
    so = L.build()                      # cross-compiles for sm_100a without a GPU
    assert so.exists()
    dll = ctypes.CDLL(str(so))          # loads without a GPU: no CUDA call at load time
    missing = [s for s in sorted(declared) if not hasattr(dll, s)]
    assert not missing, missing
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    assert dll.vc_abi_version() == L.ABI_VERSION


def test_product_never_imports_oracle():
    pkg = ROOT / "video-caption-algorithm_b200"
    for f in pkg.rglob("*.py"):
        txt = f.read_text()
        assert "oracle" not in re.sub(r'""".*?"""', "", txt, flags=re.S).replace("# oracle", ""), f
    for f in (pkg / "csrc").glob("*"):
        if f.suffix in {".cu", ".cuh", ".h"}:
            assert "oracle" not in f.read_text(), f


def test_compute_entry_points_fail_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vcb200.model import B200CaptionModel
    with pytest.raises(L.VcError):
        B200CaptionModel({}, "cuda:0", vit_heads=12, gpt_heads=12)


def test_synthetic_state_dict_layouts_agree():
    tv = synthetic.make_state_dict("tiny", seed=3, layout="torchvision")
    tm = synthetic.make_state_dict("tiny", seed=3, layout="timm")
    assert torch.equal(tv["encoder.backbone.model.encoder.layers.encoder_layer_1.mlp.3.weight"],
                       tm["encoder.backbone.blocks.1.mlp.fc2.weight"])
    assert torch.equal(tv["encoder.backbone.model.class_token"], tm["encoder.backbone.cls_token"])
    again = synthetic.make_state_dict("tiny", seed=3)
    assert all(torch.equal(tv[k], again[k]) for k in tv)
    assert tv["decoder.model.lm_head.weight"] is tv["decoder.model.transformer.wte.weight"]


def test_synthetic_frames_are_reproducible_per_video_index():
    a = synthetic.make_batch_u8(5, 3, 4)
    b = synthetic.make_batch_u8(6, 1, 4)
    assert a.dtype == torch.uint8 and a.shape == (3, 4, 224, 224, 3)
    assert torch.equal(a[1], b[0])
    assert not torch.equal(a[0], a[1])


def test_microbatcher_groups_requests_by_shape_in_arrival_order():
    from vcb200.serving import group_requests
    a, b = (8, 224, 224, 3), (8, 180, 240, 3)
    assert group_requests([a, a, b, a, b], max_batch=2) == [[0, 1], [2, 4], [3]]
    assert group_requests([], max_batch=4) == []
    assert group_requests([a] * 5, max_batch=64) == [[0, 1, 2, 3, 4]]


def test_checkpoint_reader_unwraps_model_state(tmp_path):
    import torch
    from vcb200.engine import read_checkpoint
    sd = {"a.weight": torch.randn(3, 2), "a.bias": torch.zeros(3)}
    torch.save(sd, tmp_path / "raw.pt")
    torch.save({"model_state": sd, "epoch": 3}, tmp_path / "wrapped.pt")
    for name in ("raw.pt", "wrapped.pt"):
        got = read_checkpoint(tmp_path / name)
        assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd)
    torch.save({"epoch": 3}, tmp_path / "bad.pt")
    with pytest.raises(ValueError):
        read_checkpoint(tmp_path / "bad.pt")


def test_resize_coefficient_tables_match_oracle_rule():
    """The product's host-side Pillow coefficient tables (resample.py) against the oracle's independent restatement, for
    shrinking, growing and identity sizes: same windows, same 22-bit weights."""
    from vcb200.resample import pillow_bilinear_coeffs
    from oracle import vc_oracle as O
    for in_size, out_size in [(480, 224), (360, 224), (1280, 224), (100, 224), (224, 224), (225, 224), (37, 224)]:
        kk, bounds, ks = pillow_bilinear_coeffs(in_size, out_size)
        wts, bnd = O.pillow_bilinear_coeffs(in_size, out_size)
        assert kk.shape == (out_size, ks) and len(wts) == out_size
        for i in range(out_size):
            assert tuple(bounds[i]) == tuple(bnd[i])
            assert kk[i].tolist() == wts[i]
            assert abs(int(kk[i].sum()) - (1 << 22)) <= ks          # normalised weights sum to 1.0 in 22-bit fixed point


def test_layernorm_folding_identity():
    """packing.fold_layernorm: LN(x) W^T + b == rstd * (x W'^T - mean * cs) + b' (what the decode chain and the encoder's
    folded GEMM epilogues compute), checked in fp64 against the unfused form on the bf16-rounded folded weights."""
    from vcb200.packing import fold_layernorm
    g = torch.Generator().manual_seed(0)
    H, N, M = 768, 96, 7
    x = torch.randn(M, H, generator=g).double() * 3 + 0.7
    w = torch.randn(N, H, generator=g) * 0.05
    b = torch.randn(N, generator=g) * 0.1
    gamma = 1 + 0.2 * torch.randn(H, generator=g)
    beta = 0.1 * torch.randn(H, generator=g)
    wf, cs, b2 = fold_layernorm(w, b, gamma, beta)
    assert wf.dtype == torch.bfloat16 and cs.dtype == torch.float32 and b2.dtype == torch.float32
    mean = x.mean(1, keepdim=True)
    rstd = 1.0 / torch.sqrt(x.var(1, unbiased=False, keepdim=True) + 1e-5)
    folded = rstd * (x @ wf.double().t() - mean * cs.double()) + b2.double()
    # unfused with the same rounded weights: gamma (.) W -> wf, beta W^T exact
    ref = ((x - mean) * rstd) @ wf.double().t() + (w.double() @ beta.double() + b.double())
    assert (folded - ref).abs().max().item() < 1e-4          # fp32 storage of cs / b'
    # and against the textbook form with unrounded weights: only the bf16 rounding of W' separates them
    full = torch.nn.functional.layer_norm(x, (H,), gamma.double(), beta.double(), 1e-5) @ w.double().t() + b.double()
    assert (folded - full).abs().max().item() < 0.05


def test_microbatcher_worker_failure_fails_every_request():
    """serving.MicroBatcher: if the worker thread dies (here: the pipeline cannot be built) every pending request's Future gets the
    exception and later submits are refused — nothing hangs.  Also: pinned staging is keyed by the frame shape, not the group size."""
    from vcb200.serving import MicroBatcher, _PinnedRing

    class _Boom:
        def pipeline(self, **kw):
            raise RuntimeError("boom")

    mb = MicroBatcher(_Boom(), max_batch=4)
    vid = torch.zeros(2, 8, 8, 3, dtype=torch.uint8)
    try:
        fut = mb.submit(vid)
        with pytest.raises(RuntimeError):
            fut.result(timeout=30)
    except RuntimeError:
        pass                                   # the worker had already failed: submit itself refuses
    mb._thread.join(timeout=30)
    assert isinstance(mb.error, RuntimeError)
    with pytest.raises(RuntimeError):
        mb.submit(vid)
    if torch.cuda.is_available():
        ring = _PinnedRing((2, 8, 8, 3), max_batch=4, depth=3)
        a, b = ring.next(1), ring.next(4)
        assert a.shape == (1, 2, 8, 8, 3) and b.shape == (4, 2, 8, 8, 3) and len(ring.bufs) == 3 and a.is_pinned()


def test_frame_decode_pool_equals_single_thread_pil(tmp_path):
    """frames.FrameDecodePool: JPEG files decoded by the thread pool are byte-identical to the reference's own per-file
    `Image.open(p).convert("RGB")` (frame_loader.py:42-45), for paths and for encoded bytes, with the reference's sampling."""
    import numpy as np
    from PIL import Image
    from vcb200.frames import FrameDecodePool, list_sampled_frames
    dirs = []
    for v in range(3):
        d = tmp_path / f"clip{v}"
        d.mkdir()
        fr = synthetic.make_frames_u8(20 + v, num_frames=9, size=96).numpy()
        for i in range(9):
            Image.fromarray(fr[i]).save(d / f"frame_{i:06d}.jpg", quality=90)
        dirs.append(d)
    with FrameDecodePool(workers=4, pin=False) as pool:
        got = pool.decode_dirs(dirs, num_frames=4).clone()
        picks = [list_sampled_frames(d, 4) for d in dirs]
        assert [p.name for p in picks[0]] == ["frame_000000.jpg", "frame_000002.jpg", "frame_000004.jpg", "frame_000006.jpg"]
        want = np.stack([np.stack([np.asarray(Image.open(p).convert("RGB")) for p in ps]) for ps in picks])
        assert got.shape == (3, 4, 96, 96, 3) and np.array_equal(got.numpy(), want)
        as_bytes = [[p.read_bytes() for p in ps] for ps in picks]
        assert np.array_equal(pool.decode(as_bytes).numpy(), want)
        with pytest.raises(ValueError):
            pool.decode([picks[0], picks[1][:3]])


def test_epilogue_gelu_form_matches_erf_gelu():
    """The GEMM epilogues evaluate nn.GELU() (erf form, torchvision MLPBlock / video_encoder.py:82-103) as
    x * sigmoid(x (c1 + c3 x^2 + c5 x^4)) with x^2 clamped at 52.6 (csrc/vc_common.cuh: gelu_erf_fast).  The constants in the
    kernel carry -log2(e); this restates the formula in fp32 and checks it against the erf GELU: max abs error <= 3e-5
    (the tanh formula is 4.7e-4 away), exact limits for large |x|."""
    import re
    src = (ROOT / "video-caption-algorithm_b200" / "csrc" / "vc_common.cuh").read_text()
    body = src[src.index("float gelu_erf_fast(float x)"):]
    body = body[:body.index("\n}\n")]
    consts = [float(v) for v in re.findall(r"(-?\d+\.\d+(?:e-?\d+)?)f", body)]
    clamp, k5, k3, k1 = consts[0], consts[1], consts[2], consts[3]
    assert clamp == 52.6 and k5 > 0 and k3 < 0 and k1 < 0, consts
    x = torch.linspace(-30, 30, 600001, dtype=torch.float32)
    x2 = torch.clamp(x * x, max=clamp)
    q = (x2 * k5 + k3) * x2 + k1
    got = x / (1.0 + torch.exp2(x * q))
    ref = torch.nn.functional.gelu(x.double()).float()
    assert (got - ref).abs().max().item() <= 3e-5
    assert not torch.isnan(got).any()
    assert got[0].item() == 0.0 and got[-1].item() == 30.0
