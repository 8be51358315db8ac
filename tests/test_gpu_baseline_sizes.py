"""Parity of the CUDA path AT THE SIZES BASELINE.json names (the tiny cuts of test_gpu_path.py exercise the logic; these
exercise the grids the bench actually launches), with the CPU oracle as the checker:

  cfg2  64 videos x 16 frames, ViT-B/16 + GPT-2 small at full depth, through CaptionPipeline with 4 batches per decode chain
        (256-row chain = the path bench.py times) and through the single-batch call (64-row chain);
  cfg4  64 videos x 5 beams x 30 tokens (320 rows);
  cfg5  ViT-L/14 + GPT-2 medium at full depth, 32 frames per clip;
  the eos / finished branch of the greedy loop, against the reference's own loop (tests/golden/eos_tiny.npz) and the oracle.

Tolerances as in test_gpu_path.py (anchored to the reference's own bf16-vs-fp32 error)."""
import ctypes as C

import numpy as np
import pytest
import torch

import vcb200  # noqa: F401
from vcb200 import lib as L
from vcb200 import synthetic
from vcb200.model import B200CaptionModel
from oracle import vc_oracle as O
from oracle import eos_fixture as EF

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FEAT_MAXABS, FEAT_COS = 0.02, 0.9999
LOGIT_MAXABS, LOGIT_COS = 0.06, 0.9995
TF_AGREEMENT = 0.93


def _model(arch, seed=1234, sd=None, **kw):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    a = synthetic.ARCHS[arch]
    sd = sd if sd is not None else synthetic.make_state_dict(a, seed=seed)
    return a, sd, B200CaptionModel(sd, DEV, vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, **kw)


def _cos_min(a, b):
    return torch.nn.functional.cosine_similarity(a.flatten(1).float(), b.flatten(1).float(), dim=-1).min().item()


# ------------------------------------------------------------------------------------------------------------------ eos / finished
def _decisive_ids_equal(ids, ids_ref, lens_ref, top2, thr=4 * LOGIT_MAXABS, rows=None, stop_at_tie=True):
    """ids must match up to the first step whose top-2 margin in the reference is not decisive (a free-running row may fork
    there; teacher-forced rows never fork, so `stop_at_tie=False` compares every decisive step); slots past the length hold eos."""
    n_rows, n_new = ids_ref.shape
    for r in (range(n_rows) if rows is None else rows):
        n = int(lens_ref[r])
        assert ids[r, n:].eq(EF.EOS).all(), r
        for s in range(n):
            if top2[s, r] > thr:
                assert int(ids[r, s]) == int(ids_ref[r, s]), (r, s)
            elif stop_at_tie:
                break


@pytest.mark.parametrize("n_rows", [32, 80])
def test_eos_branch_free_running_against_reference_fixture_and_oracle(golden_dir, n_rows):
    """benchmark_baseline.py:212-224.  32 rows: the reference's own loop (fixture) — prefill on the split-K chain (160 rows),
    steps on decode_chain.cu (<= 64 rows).  80 rows: every forward on the split-K chain + greedy_select_kernel, vs the oracle."""
    g = np.load(golden_dir / "eos_tiny.npz")
    a = synthetic.ARCHS["tiny"]
    sd = EF.doctor(synthetic.make_state_dict(a, seed=int(g["seed"])), a.gpt_dim)
    _, _, m = _model("tiny", sd=sd)
    n_new = int(g["max_new_tokens"])
    prefix = EF.prefixes(n_rows, a.prefix_len, a.gpt_dim)
    if n_rows == int(g["n_rows"]):
        ids_ref, lens_ref = torch.from_numpy(g["ids"]), torch.from_numpy(g["lengths"])
        margin, top2 = torch.from_numpy(g["eos_margin"]), torch.from_numpy(g["top2_margin"])
    else:
        ids_ref, lens_ref, lg = O.greedy_decode(sd, prefix, torch.tensor([[EF.PROMPT]]), n_new, heads=a.gpt_heads, keep_logits=True)
        Lg = torch.stack(lg, 0)
        margin = EF.eos_margin(Lg)
        t2 = Lg.topk(2, dim=-1).values
        top2 = t2[..., 0] - t2[..., 1]
    # rows whose eos decision is never a near-tie while they are live (5 x the logit tolerance): rows are independent, so the
    # whole batch runs and these rows are compared — they must include rows that stop at step 0 and rows that never stop
    rows = [r for r in range(n_rows) if margin[: int(lens_ref[r]), r].abs().min().item() > 5 * LOGIT_MAXABS]
    assert len(rows) >= n_rows * 3 // 4 and {int(lens_ref[r]) for r in rows} >= {1, n_new}
    ids, lens, _ = m.greedy_ids(prefix.to(DEV), [EF.PROMPT], n_new)
    torch.cuda.synchronize()
    assert [int(lens[r]) for r in rows] == [int(lens_ref[r]) for r in rows]
    _decisive_ids_equal(ids.cpu(), ids_ref, lens_ref, top2, rows=rows)
    if n_rows == int(g["n_rows"]):
        # every row of the batch finishes at step 0: lengths 1, all later slots eos (the reference loop breaks after one forward)
        first = [r for r in g["first_rows"].tolist() if r in rows]
        assert len(first) >= 5
        ids1, lens1, _ = m.greedy_ids(prefix[first].to(DEV), [EF.PROMPT], n_new)
        torch.cuda.synchronize()
        assert lens1.cpu().tolist() == [1] * len(first)
        assert ids1.cpu().tolist() == [g["ids"][r].tolist() for r in first]


def test_eos_branch_finishes_at_chosen_steps_under_teacher_forcing():
    """Rows made to finish at steps 1, 3 and 7 by feeding the trigger token (oracle/eos_fixture.py), rows that finish at step 0
    and rows that never finish, in one batch: ids and lengths equal the oracle's loop bit for bit where its margins are decisive."""
    a = synthetic.ARCHS["tiny"]
    sd = EF.doctor(synthetic.make_state_dict(a, seed=1234), a.gpt_dim)
    _, _, m = _model("tiny", sd=sd)
    n_rows, n_new = 24, 10
    prefix = EF.prefixes(n_rows, a.prefix_len, a.gpt_dim)
    _, lens0, _ = O.greedy_decode(sd, prefix, torch.tensor([[EF.PROMPT]]), n_new, heads=a.gpt_heads)
    forced = torch.randint(100, 40000, (n_rows, n_new), generator=torch.Generator().manual_seed(3))
    live = [r for r in range(n_rows) if lens0[r] == n_new][:3]
    for r, s in zip(live, (0, 2, 6)):
        forced[r, s] = EF.TRIGGER
    ids_o, lens_o, lg = O.greedy_decode(sd, prefix, torch.tensor([[EF.PROMPT]]), n_new, heads=a.gpt_heads, forced_ids=forced, keep_logits=True)
    assert [int(lens_o[r]) for r in live] == [2, 4, 8] and 1 in lens_o.tolist() and n_new in lens_o.tolist()
    Lg = torch.stack(lg, 0)
    t2 = Lg.topk(2, dim=-1).values
    ids, lens, _ = m.greedy_ids(prefix.to(DEV), [EF.PROMPT], n_new, forced_ids=forced.to(DEV))
    torch.cuda.synchronize()
    assert lens.cpu().tolist() == lens_o.tolist()
    _decisive_ids_equal(ids.cpu(), ids_o, lens_o, t2[..., 0] - t2[..., 1], stop_at_tie=False)


def test_many_row_chain_is_repeatable():
    """The tensor-core chain (>= 128 rows) ends every residual GEMM with a "last CTA combines the row statistics" step: the
    same 256 prefixes must give the same ids and logits on every run, with and without other work on the GPU."""
    a, sd, m = _model("vit_b16_gpt2", chunk_frames=1024)
    g = torch.Generator().manual_seed(11)
    prefixes = (torch.randn(256, a.prefix_len, a.gpt_dim, generator=g) * 0.5).to(DEV)
    ids0, lens0, lg0 = m.greedy_ids(prefixes, None, 20, keep_logits=True)
    ids0, lg0 = ids0.clone(), lg0.clone()
    side = torch.cuda.Stream()
    big = torch.randn(4096, 4096, device=DEV)
    for it in range(12):
        if it % 2:
            with torch.cuda.stream(side):
                for _ in range(4):
                    big @ big
        ids, lens, lg = m.greedy_ids(prefixes, None, 20, keep_logits=True)
        torch.cuda.synchronize()
        assert torch.equal(ids, ids0) and torch.equal(lg, lg0), f"run {it} differs"


# ------------------------------------------------------------------------------------------------------------------ cfg2
@pytest.mark.parametrize("G", [4, 8])
def test_cfg2_pipeline_at_bench_size_against_oracle(G):
    """BASELINE.json configs[1]: 64 videos x 16 frames, full depth, greedy 20 tokens — the exact path bench.py times
    (CaptionPipeline, decode_group=8 -> one 512-row decode chain per 8 encoder batches; 4 -> 256 rows, the earlier default)."""
    a, sd, m = _model("vit_b16_gpt2", chunk_frames=1024)
    B, T, n_new = 64, 16, 20
    batches = [synthetic.make_batch_u8(64 * i, B, T) for i in range(G)]
    pipe = m.pipeline(max_new_tokens=n_new, decode_group=G)
    tickets = [pipe.submit(f.to(DEV)) for f in batches]
    got = [tuple(x.clone() for x in pipe.result(t)) for t in tickets]
    pipe.drain()
    # (1) features of all 64 videos of the first batch vs the oracle
    feat, prefix = m.encode_prefix(batches[0].to(DEV))
    torch.cuda.synchronize()
    feat, prefix = feat.cpu(), prefix.cpu()
    n_o = B if G == 4 else 8               # the oracle costs ~0.3 s per video: all 64 once, the 8 logit videos in the second case
    feat_o = torch.cat([O.encode(sd, O.preprocess_u8(batches[0][i:i + 8]), a.vit_heads) for i in range(0, n_o, 8)], 0)
    assert (feat[:n_o] - feat_o).abs().max().item() <= FEAT_MAXABS
    assert _cos_min(feat[:n_o], feat_o) >= FEAT_COS
    prefix_o = O.visual_prefix(sd, feat_o)
    assert _cos_min(prefix[:n_o], prefix_o) >= FEAT_COS
    # (2) the pipeline's ids are those of ONE greedy call over the G x 64 prefixes (same chain, same rows, same order)
    prefixes = torch.cat([m.encode_prefix(f.to(DEV))[1] for f in batches], 0)
    ids256, lens256, _ = m.greedy_ids(prefixes, None, n_new)
    torch.cuda.synchronize()
    ids_pipe = torch.cat([g[0] for g in got], 0)
    assert torch.equal(ids_pipe, ids256.cpu()) and torch.equal(torch.cat([g[1] for g in got], 0), lens256.cpu())
    # (3) teacher-forced logits of 8 videos against the oracle, through BOTH row-count regimes: the G x 64-row chain (rows 0..7 of
    #     the grouped call) and the 64-row chain of a single batch
    sel = list(range(8))
    ids_o, lens_o, lg_o = O.greedy_decode(sd, prefix_o[sel], torch.tensor([[50256]]), n_new, heads=a.gpt_heads, keep_logits=True)
    Lo = torch.stack(lg_o, 0)                                                  # [steps, 8, V]
    steps = Lo.shape[0]
    forced256 = torch.full((G * B, n_new), 11, dtype=torch.int32)
    forced256[sel] = ids_o.int()
    pre256 = prefixes.clone()
    pre256[sel] = prefix_o[sel].to(DEV)
    for rows in (G * B, B):
        _, _, lg = m.greedy_ids(pre256[:rows], None, n_new, forced_ids=forced256[:rows].to(DEV), keep_logits=True)
        torch.cuda.synchronize()
        lg = lg[:steps, sel].cpu()
        assert (lg - Lo).abs().max().item() <= LOGIT_MAXABS, rows
        assert _cos_min(lg.reshape(steps * 8, -1), Lo.reshape(steps * 8, -1)) >= LOGIT_COS, rows
        assert (lg.argmax(-1) == Lo.argmax(-1)).float().mean().item() >= TF_AGREEMENT, rows
    # (4) free running: wherever every top-2 margin of a video is decisive, its ids must equal the oracle's
    top2 = Lo.topk(2, dim=-1).values
    decisive = ((top2[..., 0] - top2[..., 1]).min(dim=0).values > 4 * LOGIT_MAXABS)
    for j, v in enumerate(sel):
        if decisive[j]:
            assert ids_pipe[v].tolist() == ids_o[j].tolist(), v


# ------------------------------------------------------------------------------------------------------------------ cfg4
def _gpu_step_logits_fn(m, prefix, nb):
    """Per-step logits of the CUDA forward for the ORACLE's beam bookkeeping, through the SAME launch shapes as the device-side
    beam search (prefill once per video, then B * nb rows per step), so the logits are bit-identical; the cache is replicated
    and reordered by physical copies (the HF way) instead of the slot table."""
    from vcb200.memory import KvCache
    gpt2 = m.decoder.model
    B = prefix.shape[0]
    state = {"cache": None}

    def fn(step, beam_idx, tokens):
        if step == 0:
            x = torch.cat([prefix, gpt2.transformer.wte(torch.tensor([[50256]], device=DEV).expand(B, -1))], 1)
            out = gpt2(inputs_embeds=x, past_key_values=None, s_max=64)
            c1 = out.past_key_values
            c = KvCache(m.dims["gpt_layers"], B * nb, m.dims["gpt_heads"], c1.s_max, 64, m.device)
            c.kv.copy_(c1.kv.index_select(2, torch.arange(B, device=DEV).repeat_interleave(nb)))
            c.length = c1.length
            state["cache"] = c
            torch.cuda.synchronize()
            return out.logits[:, -1, :].repeat_interleave(nb, 0).cpu()
        else:
            c = state["cache"]
            c.kv.copy_(c.kv.index_select(2, beam_idx.to(DEV)))
            out = gpt2(inputs_embeds=gpt2.transformer.wte(tokens.to(DEV)).unsqueeze(1), past_key_values=c)
        state["cache"] = out.past_key_values
        torch.cuda.synchronize()
        return out.logits[:, -1, :].cpu()
    return fn


def test_cfg4_beam5_30_tokens_batch64_equals_oracle_bookkeeping():
    """BASELINE.json configs[3]: 64 videos, 5 beams, 30 tokens (320 rows, 5 row tiles of the decode kernels), HF semantics of
    text_decoder.py:131-144.  The oracle's HF-pinned bookkeeping is driven by the CUDA forward's own logits; the device-side
    selection + slot-table cache reorder must return the same ids and lengths for all 64 videos."""
    a, sd, m = _model("vit_b16_gpt2")
    B, nb, mx = 64, 5, 30
    g = torch.Generator().manual_seed(4)
    prefix = (torch.randn(B, a.prefix_len, a.gpt_dim, generator=g) * 0.3).to(DEV)
    from vcb200.decoding import hf_generate_ids
    ids, lens = hf_generate_ids(m, prefix, [50256], max_new_tokens=mx, num_beams=nb, no_repeat_ngram_size=3, repetition_penalty=1.1,
                                min_new_tokens=8)
    torch.cuda.synchronize()
    ids_o, len_o = O.beam_search(sd, prefix.cpu(), torch.tensor([[50256]]), num_beams=nb, max_new_tokens=mx, heads=a.gpt_heads,
                                 no_repeat_ngram_size=3, repetition_penalty=1.1, min_new_tokens=8,
                                 step_logits_fn=_gpu_step_logits_fn(m, prefix, nb))
    same = [ids[b].cpu().tolist() == ids_o[b].tolist() and int(lens[b]) == int(len_o[b]) for b in range(B)]
    assert all(same), [b for b in range(B) if not same[b]]
    assert all(8 <= int(n) <= mx for n in lens.tolist())


def test_beam5_30_against_reference_golden(golden_dir):
    """`beam5_30` of both fixtures (the reference's decoder.generate(num_beams=5, max_new_tokens=30) in fp32)."""
    from vcb200.decoding import hf_generate_ids
    for arch in ("tiny", "vit_b16_gpt2"):
        g = np.load(golden_dir / f"path_{arch}.npz")
        a, sd, m = _model(arch, int(g["seed"]))
        ref = torch.from_numpy(g["beam5_30"])
        ids, lens = hf_generate_ids(m, torch.from_numpy(g["prefix"]).to(DEV), [50256], max_new_tokens=30, num_beams=5,
                                    no_repeat_ngram_size=3, repetition_penalty=1.1, min_new_tokens=8)
        torch.cuda.synchronize()
        got = ids.cpu()[:, : ref.shape[1]]
        # bf16 beams vs the fp32 reference may fork at a near-tie; before the fork they are identical
        fork = [next((i for i in range(ref.shape[1]) if int(got[b, i]) != int(ref[b, i])), ref.shape[1]) for b in range(ref.shape[0])]
        assert min(fork) >= 4, (arch, fork)
        assert np.mean(fork) >= 0.5 * ref.shape[1], (arch, fork)


# ------------------------------------------------------------------------------------------------------------------ cfg5
def test_cfg5_vit_l14_gpt2_medium_full_depth_against_oracle():
    """BASELINE.json configs[4] at full depth (24 + 24 layers, width 1024, 257 tokens, 32 frames per clip), 2 videos."""
    a, sd, m = _model("vit_l14_gpt2m", seed=5, chunk_frames=512)
    frames = synthetic.make_batch_u8(40, 2, 32)
    n_new = 6
    ids_o, len_o, feat_o, prefix_o = O.caption_ids(sd, frames, vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, max_new_tokens=n_new)
    feat, prefix = m.encode_prefix(frames.to(DEV))
    torch.cuda.synchronize()
    assert (feat.cpu() - feat_o).abs().max().item() <= FEAT_MAXABS
    assert _cos_min(feat.cpu(), feat_o) >= FEAT_COS
    _, _, logits_o = O.greedy_decode(sd, prefix_o, torch.tensor([[50256]]), n_new, heads=a.gpt_heads, forced_ids=ids_o, keep_logits=True)
    _, _, logits = m.greedy_ids(prefix_o.to(DEV), None, n_new, forced_ids=ids_o.to(DEV), keep_logits=True)
    torch.cuda.synchronize()
    Lo = torch.stack(logits_o, 0)
    lg = logits.cpu()[: Lo.shape[0]]
    assert (lg - Lo).abs().max().item() <= LOGIT_MAXABS
    assert _cos_min(lg.flatten(0, 1), Lo.flatten(0, 1)) >= LOGIT_COS
    assert (lg.argmax(-1) == Lo.argmax(-1)).float().mean().item() >= TF_AGREEMENT


# ------------------------------------------------------------------------------------------------------------------ NCCL gather
def test_nccl_gather_world2_block_order():
    """§8e: the path's only exchange.  Two ranks on two GPUs of this box caption their own shard and all-gather the packed ids:
    block r of the result on every rank equals rank r's own ids.  Skipped on a one-GPU box (the gloo test covers the logic)."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import os, subprocess, sys, tempfile
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    code = (
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {str(root)!r})\n"
        "import vcb200; from vcb200 import synthetic; from vcb200.model import B200CaptionModel; from vcb200.sharding import IdGatherer\n"
        "r = int(os.environ['RANK']); torch.cuda.set_device(r); dev = torch.device('cuda', r)\n"
        "dist.init_process_group('nccl', device_id=dev)\n"
        "a = synthetic.ARCHS['tiny']; m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), dev, vit_heads=a.vit_heads, gpt_heads=a.gpt_heads)\n"
        "ids, lens = m.caption_ids(synthetic.make_batch_u8(3 * r, 3, 2).to(dev), max_new_tokens=5)\n"
        "g = IdGatherer(3, 5, 2, dev); out = g.gather(ids, lens); torch.cuda.synchronize()\n"
        "assert g.check_own_block(out, ids, lens, r)\n"
        "both = [torch.empty_like(out) for _ in range(2)]; dist.all_gather(both, out); assert torch.equal(both[0], both[1])\n"
        "dist.destroy_process_group(); print('ok', r)\n")
    with tempfile.TemporaryDirectory() as td:
        f = Path(td) / "w.py"
        f.write_text(code)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                            "--master-port", "29533", str(f)], capture_output=True, text=True, timeout=600, env=dict(os.environ))
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
