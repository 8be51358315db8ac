"""HF-`generate` semantics for `decoder.generate` (src/models/text_decoder.py:131-144) on the
B200 path: greedy with logits processors (num_beams == 1) and beam search (num_beams > 1).

Heavy work per step runs on the device through the C ABI: the GPT-2 forward over all running
rows, log-softmax + RepetitionPenalty / NoRepeatNGram / MinNewTokens processors + running
scores + the top-2*num_beams continuation search (`vc_beam_step`), and the KV-cache beam
reorder as a slot-table update (`vc_beam_reorder`) instead of HF's per-layer index_select.
The bookkeeping over B x 2*num_beams candidates per step (transformers
`_get_running_beams_for_next_iteration`, `_update_finished_beams`, `_check_early_stop_heuristic`,
SURVEY.md A.4) is one small kernel per step (`vc_beam_update`), so the whole search is a fixed
sequence of launches captured as a CUDA graph: no torch op and no host round trip per step (HF's own
loop synchronises every step).

The prefill runs ONCE per video (B rows); the num_beams-fold replication HF performs is
expressed through the slot table (every beam of a video reads the prompt positions from the
video's row), and the first selection step merges over one row per video because beams
1..nb-1 start at -1e9 and cannot reach the top 2*num_beams.
"""
from __future__ import annotations

import ctypes as C
from typing import List

import torch

from . import lib as L
from .memory import KvCache

EOS = 50256
NEG = -1.0e9


def beam_search_ids(m, prefix: torch.Tensor, prompt_ids: List[int], *, max_new_tokens: int, num_beams: int,
                    no_repeat_ngram_size: int = 3, repetition_penalty: float = 1.1, min_new_tokens: int = 8,
                    length_penalty: float = 1.0, eos: int = EOS, do_sample: bool = False, temperature: float = 1.0, top_p: float = 1.0,
                    top_k: int = 50, generator=None, use_graph: bool = True):
    if num_beams == 1:
        return _greedy_with_processors(m, prefix, prompt_ids, max_new_tokens, no_repeat_ngram_size, repetition_penalty,
                                       min_new_tokens, eos, do_sample=do_sample, temperature=temperature, top_p=top_p, top_k=top_k,
                                       generator=generator)
    with torch.cuda.device(m.device):
        return _beam_search_device(m, prefix, prompt_ids, max_new_tokens, num_beams, int(no_repeat_ngram_size), float(repetition_penalty),
                                   int(min_new_tokens), float(length_penalty), int(eos), use_graph)


def _beam_state(m, B: int, P: int, Lp: int, nb: int, max_new: int, eos: int):
    """Every buffer of one beam-search shape, allocated once and kept (m._graphs): KV cache with two slot tables, workspace,
    logits, candidate buffers, the VcBeamState arrays, static input / output tensors and the captured graphs."""
    d, dev = m.dims, m.device
    key = ("beam", B, P, Lp, nb, max_new, eos)
    st = m._graphs.get(key)
    if st is not None:
        return st
    lib = L.load()
    H, ld = d["gpt_dim"], d["vocab_pad"]
    n_rows, K, L0 = B * nb, 2 * nb, P + Lp
    s_max = L0 + max_new
    cache = KvCache(d["gpt_layers"], n_rows, d["gpt_heads"], s_max, 64, dev, with_slots=True)
    i32 = lambda *shape: torch.zeros(*shape, device=dev, dtype=torch.int32)
    f32 = lambda *shape: torch.zeros(*shape, device=dev, dtype=torch.float32)
    st = dict(cache=cache, slot_a=cache.slot, slot_b=cache.slot.clone(), slot_init=None,
              ws=torch.empty(lib.vc_gpt_workspace_bytes(C.byref(m.packed.gpt), n_rows, max(n_rows, B * L0)), device=dev, dtype=torch.uint8),
              logits=f32(n_rows, ld), embeds=f32(n_rows, H), x0=f32(B, L0, H), prompt=i32(max(Lp, 1)), tok_emb=f32(max(Lp, 1), H),
              cand_score=f32(n_rows, K), cand_tok=i32(n_rows, K), top_score=f32(B, K), top_idx=i32(B, K), zeros=f32(n_rows),
              running_scores=f32(B, nb), running_seqs=i32(n_rows, max_new), fin_seqs=i32(B, nb, max_new), fin_scores=f32(B, nb),
              fin_done=i32(B, nb), fin_len=i32(B, nb), unsatisfied=i32(B), flags=i32(max_new + 1, 2), stopped=i32(1),
              src_rows=i32(n_rows), next_tok=i32(n_rows), ids=i32(B, max_new), lens=i32(B), graphs={}, prompt_host=None)
    # every beam row of video b reads the prompt positions from physical row b (prefilled once); decode positions follow the beam
    init = torch.arange(n_rows, device=dev, dtype=torch.int32).view(n_rows, 1).repeat(1, s_max)
    init[:, :L0] = (torch.arange(n_rows, device=dev, dtype=torch.int32) // nb).view(n_rows, 1)
    st["slot_init"] = init.contiguous()
    bs = L.VcBeamState()
    bs.B, bs.nb, bs.max_len, bs.eos = B, nb, max_new, eos
    for name in ("running_scores", "running_seqs", "fin_seqs", "fin_scores", "fin_done", "fin_len", "unsatisfied", "flags", "stopped",
                 "src_rows", "next_tok"):
        setattr(bs, name, st[name].data_ptr())
    st["bs"] = bs
    m._graphs[key] = st
    return st


def _beam_search_device(m, prefix, prompt_ids, max_new_tokens, nb, ngram, rep_penalty, min_new, length_penalty, eos, use_graph):
    """transformers `_beam_search` as a FIXED sequence of launches: prefill once per video, then per step
    vc_beam_step (log-softmax, processors, top 2*nb continuations) -> vc_beam_update (running / finished hypotheses, early-stop
    heuristic) -> vc_beam_reorder (slot-table update) -> wte gather -> forward over B*nb rows.  No torch op and no host read inside
    the loop, so the whole search is captured once per shape as a CUDA graph and replayed.  HF leaves its loop as soon as the
    early-stop heuristic is satisfied; here the remaining steps still run but the finished pool is frozen from that step on
    (VcBeamState.stopped), so the result is the same."""
    d, dev = m.dims, m.device
    lib = L.load()
    gpt = C.byref(m.packed.gpt)
    B, P, H = prefix.shape
    V, ld = d["vocab"], d["vocab_pad"]
    K, n_rows, Lp = 2 * nb, B * nb, len(prompt_ids)
    L0 = P + Lp
    s_max = L0 + max_new_tokens
    st = _beam_state(m, B, P, Lp, nb, max_new_tokens, eos)
    if st["prompt_host"] != tuple(prompt_ids):
        st["prompt"][:Lp].copy_(torch.tensor(prompt_ids, dtype=torch.int32))
        st["prompt_host"] = tuple(prompt_ids)
    st["x0"][:, :P].copy_(prefix.to(device=dev, dtype=torch.float32))
    stream = lambda: L.current_stream(dev)
    bs = C.byref(st["bs"])

    def enqueue():
        cache = st["cache"]
        slot_a, slot_b = st["slot_a"], st["slot_b"]
        # BOTH tables start from the initial mapping on every run: a reorder only rewrites positions below `past`, so the entry of
        # the position a step is about to write must still say "own row" — it would hold the previous run's beam otherwise
        slot_a.copy_(st["slot_init"])
        slot_b.copy_(st["slot_init"])
        cache.c.slot = slot_a.data_ptr()
        L.check(lib.vc_beam_init(bs, stream()))
        L.check(lib.vc_gpt2_embed_tokens(gpt, st["prompt"].data_ptr(), Lp, st["tok_emb"].data_ptr(), stream()))
        st["x0"][:, P:].copy_(st["tok_emb"][:Lp].unsqueeze(0).expand(B, -1, -1))
        L.check(lib.vc_gpt2_forward(gpt, st["x0"].data_ptr(), B, L0, 0, C.byref(cache.c), st["ws"].data_ptr(), st["ws"].numel(),
                                    st["logits"].data_ptr(), 0, stream()))
        for cur_len in range(max_new_tokens):
            first = cur_len == 0
            rows, per_item = (B, 1) if first else (n_rows, nb)
            running = st["zeros"] if first else st["running_scores"]
            L.check(lib.vc_beam_step(st["logits"].data_ptr(), ld, V, rows, per_item, st["running_seqs"].data_ptr(), max_new_tokens, cur_len,
                                     running.data_ptr(), rep_penalty, ngram, min_new, eos, 0, K, st["cand_score"].data_ptr(),
                                     st["cand_tok"].data_ptr(), st["top_score"].data_ptr(), st["top_idx"].data_ptr(), stream()))
            L.check(lib.vc_beam_update(bs, st["top_score"].data_ptr(), st["top_idx"].data_ptr(), V, cur_len, length_penalty, stream()))
            if cur_len + 1 >= max_new_tokens:
                break
            past = L0 + cur_len
            if not first:
                # positions written by decode steps follow their beam; prompt positions keep the per-video mapping
                L.check(lib.vc_beam_reorder(slot_a.data_ptr(), slot_b.data_ptr(), st["src_rows"].data_ptr(), n_rows, s_max, past, stream()))
                slot_a, slot_b = slot_b, slot_a
                cache.c.slot = slot_a.data_ptr()
            L.check(lib.vc_gpt2_embed_tokens(gpt, st["next_tok"].data_ptr(), n_rows, st["embeds"].data_ptr(), stream()))
            L.check(lib.vc_gpt2_forward(gpt, st["embeds"].data_ptr(), n_rows, 1, past, C.byref(cache.c), st["ws"].data_ptr(), st["ws"].numel(),
                                        st["logits"].data_ptr(), 0, stream()))
        L.check(lib.vc_beam_finalize(bs, st["ids"].data_ptr(), st["lens"].data_ptr(), stream()))

    gkey = (ngram, rep_penalty, min_new, length_penalty)
    if not use_graph:
        enqueue()
    elif gkey not in st["graphs"]:
        enqueue()                                   # eager warm-up: func attributes are set outside the capture
        torch.cuda.current_stream(dev).synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=torch.cuda.Stream(device=dev)):
            enqueue()
        st["graphs"][gkey] = g
        g.replay()
    else:
        st["graphs"][gkey].replay()
    return st["ids"].clone(), st["lens"].clone()           # the graph's static outputs are reused by the next call of this shape


def _processed_scores(logits: torch.Tensor, seqs: torch.Tensor, cur_len: int, V: int, ngram: int, rep_penalty: float, min_new: int,
                      eos: int) -> torch.Tensor:
    """transformers RepetitionPenalty -> NoRepeatNGram -> MinNewTokensLength processors on full rows, as device torch ops
    (the sampling path needs the whole distribution; the argmax / beam paths use the fused `vc_beam_step` kernel).
    `seqs` holds the generated tokens only: `generate(inputs_embeds=...)` starts from empty input_ids."""
    scores = logits[:, :V].clone()
    if cur_len > 0:
        prev = seqs[:, :cur_len].to(torch.int64)
        if rep_penalty != 1.0:
            sc = torch.gather(scores, 1, prev)
            pen = torch.full_like(sc, rep_penalty)          # tensor / tensor is an IEEE division; tensor / python-scalar on CUDA
            sc = torch.where(sc < 0, sc * pen, sc / pen)    # multiplies by the rounded reciprocal (1 ulp off the CPU result)
            scores.scatter_(1, prev, sc)
        if ngram > 0 and cur_len + 1 >= ngram:
            # ban every token that would complete an n-gram already present: windows whose first ngram-1 tokens equal the tail
            n_win = cur_len - ngram + 1
            if n_win > 0:
                match = torch.ones(prev.shape[0], n_win, dtype=torch.bool, device=prev.device)
                for j in range(ngram - 1):
                    match &= prev[:, j:j + n_win] == prev[:, cur_len - (ngram - 1) + j].unsqueeze(1)
                banned = prev[:, ngram - 1:ngram - 1 + n_win]
                # a token can close a matching and a non-matching window: min-reduce so that one match is enough
                val = torch.where(match, torch.full((), float("-inf"), device=scores.device), torch.full((), float("inf"), device=scores.device))
                scores.scatter_reduce_(1, banned, val, reduce="amin", include_self=True)
    if cur_len < min_new:
        scores[:, eos] = float("-inf")
    return scores


def _warped_scores(scores: torch.Tensor, temperature: float, top_p: float, top_k: int = 50) -> torch.Tensor:
    """transformers TemperatureLogitsWarper -> TopKLogitsWarper -> TopPLogitsWarper, the order `_get_logits_processor` builds them
    in.  The reference calls `generate(do_sample=True, temperature, top_p)` without a top_k, so GenerationConfig's default
    top_k = 50 applies (text_decoder.py:131-144): everything below the 50th largest score is masked before the nucleus is cut.
    Top-p (min_tokens_to_keep=1): ascending sort, drop the tokens whose cumulative probability stays <= 1 - top_p, always keep
    the most probable one."""
    scores = scores / torch.full_like(scores[:, :1], float(temperature))
    if top_k and top_k > 0:
        k = min(int(top_k), scores.shape[-1])
        kth = torch.topk(scores, k, dim=-1).values[:, -1:]
        scores = scores.masked_fill(scores < kth, float("-inf"))
    if top_p < 1.0:
        srt, idx = torch.sort(scores, descending=False, dim=-1)
        remove = srt.softmax(dim=-1).cumsum(dim=-1) <= (1.0 - float(top_p))
        remove[:, -1:] = False
        scores = scores.masked_fill(remove.scatter(1, idx, remove), float("-inf"))
    return scores


def _greedy_with_processors(m, prefix, prompt_ids, max_new_tokens, ngram, rep_penalty, min_new, eos, *, do_sample: bool = False,
                            temperature: float = 1.0, top_p: float = 1.0, top_k: int = 50, generator=None):
    """HF `_sample`: processors on the raw last-position logits, then argmax (do_sample=False) or temperature / top-p warpers,
    softmax and `torch.multinomial` (do_sample=True: text_decoder.py:137, presets "natural" / "safe_sample"); finished rows
    emit pad(=eos); stop when every row has produced eos or at max_new_tokens.  All state stays on the device: the host reads
    the all-finished flag every fourth step.  Sampling draws from torch's CUDA generator, so it is reproducible under
    `torch.manual_seed` but is not comparable token-for-token with another implementation (excluded from parity)."""
    d = m.dims
    dev = m.device
    lib = L.load()
    gpt = C.byref(m.packed.gpt)
    B, P, H = prefix.shape
    V, ld = d["vocab"], d["vocab_pad"]
    Lp = len(prompt_ids)
    L0 = P + Lp
    st = lambda: L.current_stream(dev)
    cache = KvCache(d["gpt_layers"], B, d["gpt_heads"], L0 + max_new_tokens, 64, dev)
    ws = torch.empty(lib.vc_gpt_workspace_bytes(gpt, B, B * L0), device=dev, dtype=torch.uint8)
    logits = torch.empty(B, ld, device=dev, dtype=torch.float32)
    embeds = torch.empty(B, H, device=dev, dtype=torch.float32)
    cand_score = torch.empty(B, 1, device=dev, dtype=torch.float32)
    cand_tok = torch.empty(B, 1, device=dev, dtype=torch.int32)
    top_score = torch.empty(B, 1, device=dev, dtype=torch.float32)
    top_idx = torch.empty(B, 1, device=dev, dtype=torch.int32)
    seqs_dev = torch.full((B, max_new_tokens), eos, device=dev, dtype=torch.int32)
    tok_dev = torch.empty(B, device=dev, dtype=torch.int32)
    prompt = torch.tensor(prompt_ids, device=dev, dtype=torch.int32)
    tok_emb = torch.empty(Lp, H, device=dev, dtype=torch.float32)
    L.check(lib.vc_gpt2_embed_tokens(gpt, prompt.data_ptr(), Lp, tok_emb.data_ptr(), st()))
    x0 = torch.cat([prefix.to(device=dev, dtype=torch.float32), tok_emb.unsqueeze(0).expand(B, -1, -1)], dim=1).contiguous()
    L.check(lib.vc_gpt2_forward(gpt, x0.data_ptr(), B, L0, 0, C.byref(cache.c), ws.data_ptr(), ws.numel(), logits.data_ptr(), 0, st()))
    unfinished = torch.ones(B, dtype=torch.bool, device=dev)
    lengths = torch.zeros(B, dtype=torch.int32, device=dev)
    eos_t = torch.full((B,), eos, dtype=torch.int32, device=dev)
    for cur_len in range(max_new_tokens):
        if do_sample:
            scores = _processed_scores(logits, seqs_dev, cur_len, V, int(ngram), float(rep_penalty), int(min_new), eos)
            scores = _warped_scores(scores, temperature, top_p, top_k)
            nxt = torch.multinomial(scores.softmax(dim=-1), 1, generator=generator).view(-1).to(torch.int32)
        else:
            L.check(lib.vc_beam_step(logits.data_ptr(), ld, V, B, 1, seqs_dev.data_ptr(), max_new_tokens, cur_len, 0, float(rep_penalty),
                                     int(ngram), int(min_new), eos, 1, 1, cand_score.data_ptr(), cand_tok.data_ptr(), top_score.data_ptr(),
                                     top_idx.data_ptr(), st()))
            nxt = top_idx.view(-1)
        nxt = torch.where(unfinished, nxt, eos_t)
        seqs_dev[:, cur_len] = nxt
        lengths += unfinished.to(torch.int32)
        unfinished = unfinished & (nxt != eos)
        if cur_len + 1 == max_new_tokens or ((cur_len + 1) % 4 == 0 and not bool(unfinished.any())):   # the only host sync
            break
        tok_dev.copy_(nxt)
        L.check(lib.vc_gpt2_embed_tokens(gpt, tok_dev.data_ptr(), B, embeds.data_ptr(), st()))
        L.check(lib.vc_gpt2_forward(gpt, embeds.data_ptr(), B, 1, L0 + cur_len, C.byref(cache.c), ws.data_ptr(), ws.numel(),
                                    logits.data_ptr(), 0, st()))
    return seqs_dev, lengths
