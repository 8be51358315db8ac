"""HF-`generate` semantics for `decoder.generate` (src/models/text_decoder.py:131-144) on the
B200 path: greedy with logits processors (num_beams == 1) and beam search (num_beams > 1).

Heavy work per step runs on the device through the C ABI: the GPT-2 forward over all running
rows, log-softmax + RepetitionPenalty / NoRepeatNGram / MinNewTokens processors + running
scores + the top-2*num_beams continuation search (`vc_beam_step`), and the KV-cache beam
reorder as a slot-table update (`vc_beam_reorder`) instead of HF's per-layer index_select.
The bookkeeping over B x 2*num_beams candidates per step (transformers
`_get_running_beams_for_next_iteration`, `_update_finished_beams`, `_check_early_stop_heuristic`,
SURVEY.md A.4) is a few hundred scalars: it stays in device tensors and is updated by small torch ops
on the same stream, so a step has no host round trip (HF's own loop synchronises every step); the
host reads the termination flag every fourth step.

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
                    generator=None):
    if num_beams == 1:
        return _greedy_with_processors(m, prefix, prompt_ids, max_new_tokens, no_repeat_ngram_size, repetition_penalty,
                                       min_new_tokens, eos, do_sample=do_sample, temperature=temperature, top_p=top_p, generator=generator)
    d = m.dims
    dev = m.device
    lib = L.load()
    gpt = C.byref(m.packed.gpt)
    B, P, H = prefix.shape
    nb, V, ld = num_beams, d["vocab"], d["vocab_pad"]
    K = 2 * nb
    n_rows = B * nb
    Lp = len(prompt_ids)
    L0 = P + Lp
    s_max = L0 + max_new_tokens
    st = L.current_stream

    cache = KvCache(d["gpt_layers"], n_rows, d["gpt_heads"], s_max, 64, dev, with_slots=True)
    # every beam row of video b reads the prompt positions from physical row b (prefilled once)
    slot_a = cache.slot
    slot_a[:, :L0] = (torch.arange(n_rows, device=dev, dtype=torch.int32) // nb).view(n_rows, 1)
    slot_b = slot_a.clone()
    ws = torch.empty(lib.vc_gpt_workspace_bytes(gpt, n_rows, max(n_rows, B * L0)), device=dev, dtype=torch.uint8)
    logits = torch.empty(n_rows, ld, device=dev, dtype=torch.float32)
    embeds = torch.empty(n_rows, H, device=dev, dtype=torch.float32)
    cand_score = torch.empty(n_rows, K, device=dev, dtype=torch.float32)
    cand_tok = torch.empty(n_rows, K, device=dev, dtype=torch.int32)
    top_score = torch.empty(B, K, device=dev, dtype=torch.float32)
    top_idx = torch.empty(B, K, device=dev, dtype=torch.int32)
    seqs_dev = torch.full((n_rows, max_new_tokens), eos, device=dev, dtype=torch.int32)
    run_dev = torch.zeros(n_rows, device=dev, dtype=torch.float32)
    tok_dev = torch.empty(n_rows, device=dev, dtype=torch.int32)
    src_dev = torch.empty(n_rows, device=dev, dtype=torch.int32)

    # ---- prefill: [prefix | wte(prompt)] for the B videos
    prompt = torch.tensor(prompt_ids, device=dev, dtype=torch.int32)
    tok_emb = torch.empty(Lp, H, device=dev, dtype=torch.float32)
    L.check(lib.vc_gpt2_embed_tokens(gpt, prompt.data_ptr(), Lp, tok_emb.data_ptr(), st()))
    x0 = torch.cat([prefix.to(device=dev, dtype=torch.float32), tok_emb.unsqueeze(0).expand(B, -1, -1)], dim=1).contiguous()
    L.check(lib.vc_gpt2_forward(gpt, x0.data_ptr(), B, L0, 0, C.byref(cache.c), ws.data_ptr(), ws.numel(), logits.data_ptr(), 0, st()))

    # ---- beam state ON THE DEVICE (transformers `_beam_search` variable names in comments): a few hundred scalars per
    # step, updated with small torch ops on the same stream, so a step needs no host round trip.  The host looks at the
    # termination flag every `check_every` steps; once the flag is up the finished-hypothesis state is frozen, so the
    # steps that run past HF's stopping point change nothing.
    check_every = 4
    running_scores = torch.zeros(B, nb, device=dev)
    running_scores[:, 1:] = NEG
    running_seqs = torch.full((B, nb, max_new_tokens), eos, dtype=torch.int64, device=dev)
    fin_seqs = running_seqs.clone()                                   # sequences
    fin_scores = torch.full((B, nb), NEG, device=dev)                  # beam_scores
    fin_done = torch.zeros(B, nb, dtype=torch.bool, device=dev)        # is_sent_finished
    fin_len = torch.zeros(B, nb, dtype=torch.int64, device=dev)
    unsatisfied = torch.ones(B, 1, dtype=torch.bool, device=dev)       # is_early_stop_heuristic_unsatisfied
    in_top = (torch.arange(K, device=dev).view(1, K) < nb)             # top_num_beam_mask
    row_base = (torch.arange(B, device=dev) * nb).view(B, 1)
    stopped = torch.zeros((), dtype=torch.bool, device=dev)            # HF's loop would have ended before this step
    cur_len = 0
    while True:
        first = cur_len == 0
        rows, per_item = (B, 1) if first else (n_rows, nb)
        L.check(lib.vc_beam_step(logits.data_ptr(), ld, V, rows, per_item, seqs_dev.data_ptr(), max_new_tokens, cur_len,
                                 run_dev.data_ptr(), float(repetition_penalty), int(no_repeat_ngram_size), int(min_new_tokens), eos, 0,
                                 K, cand_score.data_ptr(), cand_tok.data_ptr(), top_score.data_ptr(), top_idx.data_ptr(), st()))
        top_scores = top_score
        flat = top_idx.to(torch.int64)
        top_beam, top_tok = flat // V, flat % V
        cand_seqs = torch.gather(running_seqs, 1, top_beam.unsqueeze(-1).expand(-1, -1, max_new_tokens)).clone()
        cand_seqs[:, :, cur_len] = top_tok
        new_len = cur_len + 1
        hit_stop = (top_tok == eos) | (new_len >= max_new_tokens)
        run_rank = top_scores + hit_stop.float() * NEG
        running_scores, pick = torch.topk(run_rank, nb, dim=1)
        running_seqs = torch.gather(cand_seqs, 1, pick.unsqueeze(-1).expand(-1, -1, max_new_tokens))
        running_beam = torch.gather(top_beam, 1, pick)
        running_tok = torch.gather(top_tok, 1, pick)
        newly = hit_stop & in_top
        f_scores = top_scores / (float(new_len) ** length_penalty)
        f_scores = f_scores + (~unsatisfied).float() * NEG
        f_scores = f_scores + (~newly).float() * NEG
        merged_scores = torch.cat([fin_scores, f_scores], dim=1)
        merged_seqs = torch.cat([fin_seqs, cand_seqs], dim=1)
        merged_done = torch.cat([fin_done, newly], dim=1)
        merged_len = torch.cat([fin_len, torch.full((B, K), new_len, dtype=torch.int64, device=dev)], dim=1)
        n_scores, sel = torch.topk(merged_scores, nb, dim=1)
        n_seqs = torch.gather(merged_seqs, 1, sel.unsqueeze(-1).expand(-1, -1, max_new_tokens))
        n_done = torch.gather(merged_done, 1, sel)
        n_len = torch.gather(merged_len, 1, sel)
        # freeze the finished state once HF's loop would have stopped
        fin_scores = torch.where(stopped, fin_scores, n_scores)
        fin_seqs = torch.where(stopped, fin_seqs, n_seqs)
        fin_done = torch.where(stopped, fin_done, n_done)
        fin_len = torch.where(stopped, fin_len, n_len)
        cur_len = new_len
        best_running = running_scores[:, :1] / (float(cur_len) ** length_penalty)
        worst_fin = torch.where(fin_done, fin_scores.min(dim=1, keepdim=True).values, torch.full_like(fin_scores, NEG))
        unsatisfied = unsatisfied & (best_running > worst_fin).any(dim=-1, keepdim=True)
        stopped = stopped | ~(unsatisfied.any() & ~hit_stop.all())
        if cur_len >= max_new_tokens or (cur_len % check_every == 0 and bool(stopped)):      # the only host sync: every 4th step
            break
        # ---- next forward: reorder the cache by index, feed the chosen tokens
        src_dev.copy_((running_beam + row_base).reshape(-1))
        tok_dev.copy_(running_tok.reshape(-1))
        seqs_dev.copy_(running_seqs.reshape(n_rows, max_new_tokens))
        run_dev.copy_(running_scores.reshape(-1))
        past = L0 + cur_len - 1
        if not first:
            # positions written by decode steps follow their beam; prompt positions keep the per-video mapping
            L.check(lib.vc_beam_reorder(slot_a.data_ptr(), slot_b.data_ptr(), src_dev.data_ptr(), n_rows, s_max, past, st()))
            slot_a, slot_b = slot_b, slot_a
            cache.c.slot = slot_a.data_ptr()
        L.check(lib.vc_gpt2_embed_tokens(gpt, tok_dev.data_ptr(), n_rows, embeds.data_ptr(), st()))
        L.check(lib.vc_gpt2_forward(gpt, embeds.data_ptr(), n_rows, 1, past, C.byref(cache.c), ws.data_ptr(), ws.numel(),
                                    logits.data_ptr(), 0, st()))
    lengths = fin_len[:, 0].to(torch.int32)
    pos = torch.arange(max_new_tokens, device=dev).view(1, -1)
    ids = torch.where(pos < lengths.view(-1, 1), fin_seqs[:, 0, :], torch.full_like(fin_seqs[:, 0, :], eos)).to(torch.int32)
    return ids, lengths


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


def _warped_scores(scores: torch.Tensor, temperature: float, top_p: float) -> torch.Tensor:
    """transformers TemperatureLogitsWarper then TopPLogitsWarper (min_tokens_to_keep=1): ascending sort, drop the tokens whose
    cumulative probability stays <= 1 - top_p, always keep the most probable one."""
    scores = scores / torch.full_like(scores[:, :1], float(temperature))
    if top_p < 1.0:
        srt, idx = torch.sort(scores, descending=False, dim=-1)
        remove = srt.softmax(dim=-1).cumsum(dim=-1) <= (1.0 - float(top_p))
        remove[:, -1:] = False
        scores = scores.masked_fill(remove.scatter(1, idx, remove), float("-inf"))
    return scores


def _greedy_with_processors(m, prefix, prompt_ids, max_new_tokens, ngram, rep_penalty, min_new, eos, *, do_sample: bool = False,
                            temperature: float = 1.0, top_p: float = 1.0, generator=None):
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
    st = L.current_stream
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
            scores = _warped_scores(scores, temperature, top_p)
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
