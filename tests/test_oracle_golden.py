"""Pins oracle/vc_oracle.py against the fixtures produced by the UNMODIFIED reference
modules (oracle/pin_against_reference.py, run in the build container).  CPU only."""
import numpy as np
import pytest
import torch

import vcb200  # noqa: F401
from vcb200 import synthetic
from oracle import vc_oracle as O


def test_frame_sampling_matches_reference(golden_dir):
    g = np.load(golden_dir / "preprocess.npz")
    assert O.sample_frame_indices(int(g["n_files"]), int(g["num_frames"])) == g["picked"].tolist()
    # frame_loader.py:31-32 examples from SURVEY.md A.1
    assert O.sample_frame_indices(40, 16) == list(range(0, 32, 2))
    assert O.sample_frame_indices(10, 16) == list(range(10))
    assert O.sample_frame_indices(1, 16) == [0]


def test_preprocess_bit_exact_vs_reference(golden_dir):
    g = np.load(golden_dir / "preprocess.npz")
    u8 = torch.from_numpy(g["u8"])            # [4,32,48,3] decoded JPEG bytes
    ref = torch.from_numpy(g["ref"])          # the reference's load_video_tensor output crop
    assert torch.equal(O.preprocess_u8(u8), ref)
    lut = O.normalize_lut()
    via_lut = torch.stack([lut[c][u8[..., c].long()] for c in range(3)], dim=1)
    assert torch.equal(via_lut, ref)


@pytest.mark.parametrize("arch", ["tiny", "vit_b16_gpt2"])
def test_path_matches_reference(golden_dir, arch):
    g = np.load(golden_dir / f"path_{arch}.npz")
    a = synthetic.ARCHS[arch]
    sd = synthetic.make_state_dict(a, seed=int(g["seed"]))
    frames = synthetic.make_batch_u8(0, int(g["B"]), int(g["T"]))
    torch.set_num_threads(8)
    video = O.preprocess_u8(frames)
    feat = O.encode(sd, video, a.vit_heads)
    ref_feat = torch.from_numpy(g["feat"])
    assert (feat - ref_feat).abs().max().item() < 2e-5
    prefix = O.visual_prefix(sd, feat)
    assert (prefix - torch.from_numpy(g["prefix"])).abs().max().item() < 2e-5
    prompt = torch.tensor([[50256]])
    n_new = int(g["max_new_tokens"])
    ids, lengths, logits = O.greedy_decode(sd, prefix, prompt, n_new, heads=a.gpt_heads, keep_logits=True)
    assert ids.tolist() == g["ids"].tolist()
    assert lengths.tolist() == g["lengths"].tolist()
    L = torch.stack(logits, 0)
    stride = int(g["logits_stride"])
    assert (L[:, :, ::stride] - torch.from_numpy(g["logits_sub"])).abs().max().item() < 5e-4
    assert torch.equal(L.topk(8, dim=-1).indices, torch.from_numpy(g["logits_top_idx"]))


@pytest.mark.parametrize("arch,key,nb,mx", [("tiny", "beam3_24", 3, 24), ("tiny", "beam5_30", 5, 30),
                                            ("vit_b16_gpt2", "beam5_30", 5, 30)])
def test_beam_search_matches_hf_generate(golden_dir, arch, key, nb, mx):
    g = np.load(golden_dir / f"path_{arch}.npz")
    a = synthetic.ARCHS[arch]
    sd = synthetic.make_state_dict(a, seed=int(g["seed"]))
    prefix = torch.from_numpy(g["prefix"])
    ids, lengths = O.beam_search(sd, prefix, torch.tensor([[50256]]), num_beams=nb, max_new_tokens=mx, heads=a.gpt_heads)
    ref = g[key]                      # HF output: [B, longest], padded with eos
    for b in range(ref.shape[0]):
        n = int(lengths[b])
        assert ids[b, :n].tolist() == ref[b, :n].tolist()
        assert all(t == 50256 for t in ref[b, n:].tolist())


def test_resize_matches_pillow_golden(golden_dir):
    """oracle.resize_bilinear_u8 == the reference's transforms.Resize((224,224)) on PIL images (fixtures written by
    oracle/pin_resize_against_pillow.py), byte for byte: shrinking (antialiased), growing, mixed, near-identity sizes."""
    import numpy as np
    z = np.load(golden_dir / "resize.npz")
    names = [k for k in z.files if k.startswith("src_")]
    assert len(names) >= 5
    for k in names:
        src, want = torch.from_numpy(z[k]), torch.from_numpy(z["dst_" + k[4:]])
        got = O.resize_bilinear_u8(src, 224, 224)
        assert torch.equal(got, want), k
    # identity size is a copy
    x = torch.randint(0, 256, (2, 224, 224, 3), dtype=torch.uint8)
    assert torch.equal(O.resize_bilinear_u8(x, 224, 224), x)


def test_eos_branch_matches_reference_loop(golden_dir):
    """benchmark_baseline.py:212-224 (finished rows forced to eos, tokens appended up to and including the first eos, loop ends
    when every row has finished): the oracle's greedy loop equals the reference's own loop on weights doctored so that the
    branch fires (oracle/eos_fixture.py; fixture made by oracle/pin_against_reference.py --only-eos from the unmodified modules)."""
    from oracle import eos_fixture as EF
    g = np.load(golden_dir / "eos_tiny.npz")
    a = synthetic.ARCHS["tiny"]
    sd = EF.doctor(synthetic.make_state_dict(a, seed=int(g["seed"])), a.gpt_dim)
    n_new, n_rows = int(g["max_new_tokens"]), int(g["n_rows"])
    prefix = EF.prefixes(n_rows, a.prefix_len, a.gpt_dim)
    ids, lens, logits = O.greedy_decode(sd, prefix, torch.tensor([[EF.PROMPT]]), n_new, heads=a.gpt_heads, keep_logits=True)
    assert lens.tolist() == g["lengths"].tolist()
    assert ids.tolist() == g["ids"].tolist()
    assert set(lens.tolist()) >= {1, n_new}                                   # rows that stop at step 0 and rows that never stop
    assert all(ids[r, int(lens[r]):].eq(EF.EOS).all() for r in range(n_rows))  # later slots stay eos
    m = torch.stack([EF.eos_margin(l) for l in logits], 0)
    assert torch.allclose(m, torch.from_numpy(g["eos_margin"])[: m.shape[0]], atol=2e-3)
    # a batch in which every row finishes at step 0: the loop ends after one forward (:224)
    first = g["first_rows"].tolist()
    ids1, lens1, lg1 = O.greedy_decode(sd, prefix[first], torch.tensor([[EF.PROMPT]]), n_new, heads=a.gpt_heads, keep_logits=True)
    assert len(lg1) == int(g["first_steps_run"]) == 1
    assert ids1.tolist() == g["first_ids"].tolist() and lens1.tolist() == g["first_lengths"].tolist()
    # teacher forcing: feeding the trigger token makes the NEXT argmax eos, at any step we choose
    forced = torch.randint(100, 40000, (n_rows, n_new), generator=torch.Generator().manual_seed(3))
    live = [r for r in range(n_rows) if lens[r] == n_new][:3]
    for r, s in zip(live, (0, 2, 6)):
        forced[r, s] = EF.TRIGGER
    ids_f, lens_f, _ = O.greedy_decode(sd, prefix, torch.tensor([[EF.PROMPT]]), n_new, heads=a.gpt_heads, forced_ids=forced)
    assert [int(lens_f[r]) for r in live] == [2, 4, 8]


def test_sample_warpers_equal_transformers_processors():
    """The sampling presets (core/inference.py "natural" / "safe_sample" -> generate(do_sample=True, temperature, top_p), default
    top_k = 50): the oracle's warper chain equals the installed transformers classes in the order generate() builds them."""
    from transformers.generation.logits_process import TemperatureLogitsWarper, TopKLogitsWarper, TopPLogitsWarper
    # the reference pins transformers==4.57.1 (requirements.txt:38), whose GenerationConfig defaults to top_k = 50; the
    # transformers 5.x installed here defaults to None, so the default is stated here rather than read from the installed package
    g = torch.Generator().manual_seed(1)
    scores = torch.randn(5, 50257, generator=g) * 3
    ids = torch.zeros(5, 1, dtype=torch.long)
    for temp, top_p in ((0.9, 0.9), (0.7, 0.95), (1.3, 0.5)):
        want = scores.clone()
        for proc in (TemperatureLogitsWarper(temp), TopKLogitsWarper(top_k=50, min_tokens_to_keep=1), TopPLogitsWarper(top_p=top_p, min_tokens_to_keep=1)):
            want = proc(ids, want)
        got = O.sample_warpers(scores.clone(), temp, top_p)
        assert torch.equal(torch.isinf(got), torch.isinf(want))
        assert torch.equal(got[~torch.isinf(got)], want[~torch.isinf(want)])
        assert int((~torch.isinf(got)).sum(dim=-1).max()) <= 50


def test_postprocess_equals_reference_clean_text_and_ranker(golden_dir):
    """a13: vcb200.postprocess (clean_text, score_sentence, select_best) gives exactly what the reference's
    core/postprocessing functions gave on the pinned corpus (oracle/pin_text_against_reference.py)."""
    import json
    from vcb200 import postprocess as PP
    g = json.loads((golden_dir / "text_cleanup.json").read_text())
    for raw, want in g["clean"]:
        assert PP.clean_text(raw) == want, raw
    for s, want in g["score"]:
        assert PP.score_sentence(s) == pytest.approx(want, abs=1e-12), s
    for cands, want in g["select"]:
        key, text, score = PP.select_best([tuple(c) for c in cands])
        assert [key, text] == want[:2] and score == pytest.approx(want[2], abs=1e-12)
