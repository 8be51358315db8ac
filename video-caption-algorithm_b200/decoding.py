"""Decode drivers on top of the C ABI: the benchmark's greedy loop and HF-`generate`
semantics (processors + beam search) for `decoder.generate`.

greedy_decode  — core/scripts/benchmark_baseline.py:160-240: prefill over
  [prefix | wte(prompt)], argmax, finished rows forced to eos, tokens appended up to
  and including the first eos.  The whole loop (prefill + max_new-1 steps + all
  bookkeeping) is enqueued by ONE C call with no host sync, and replayed as a CUDA
  graph after the first call for a given shape.
hf_generate_ids — src/models/text_decoder.py:131-144 -> transformers
  `GenerationMixin.generate`: RepetitionPenalty -> NoRepeatNGram -> MinNewTokens
  processors, greedy (num_beams=1) or `_beam_search`.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import os

import torch

from . import lib as L
from .memory import KvCache

EOS = 50256


def _greedy_buffers(m, n_seq: int, P: int, Lp: int, max_new: int, keep_logits: bool):
    d = m.dims
    L0 = P + Lp
    key = ("greedy", n_seq, P, Lp, max_new, keep_logits)      # P and Lp are baked into the captured call, not only their sum
    st = m._graphs.get(key)
    if st is None:
        cache = KvCache(d["gpt_layers"], n_seq, d["gpt_heads"], L0 + max_new, 64, m.device)
        nbytes = L.load().vc_gpt_workspace_bytes(C.byref(m.packed.gpt), n_seq, n_seq * L0)
        st = dict(
            cache=cache,
            ws=torch.empty(nbytes, device=m.device, dtype=torch.uint8),
            prefix=torch.empty(n_seq, L0, d["gpt_dim"], device=m.device, dtype=torch.float32),   # only [:, :P] used
            prompt=torch.empty(max(L0, 1), device=m.device, dtype=torch.int32),
            ids=torch.empty(n_seq, max_new, device=m.device, dtype=torch.int32),
            lens=torch.empty(n_seq, device=m.device, dtype=torch.int32),
            forced=torch.empty(n_seq, max_new, device=m.device, dtype=torch.int32),
            logits=torch.empty(max_new, n_seq, d["vocab_pad"], device=m.device, dtype=torch.float32) if keep_logits else None,
            graph=None, graph_forced=None,
        )
        m._graphs[key] = st
    return st


def greedy_decode(m, prefix: torch.Tensor, prompt_ids: List[int], max_new: int, forced_ids: Optional[torch.Tensor],
                  keep_logits: bool, use_graph: bool):
    d = m.dims
    n_seq, P, H = prefix.shape
    Lp = len(prompt_ids)
    L0 = P + Lp
    st = _greedy_buffers(m, n_seq, P, Lp, max_new, keep_logits)
    pre = st["prefix"].view(-1)[: n_seq * P * H].view(n_seq, P, H)
    pre.copy_(prefix.to(device=m.device, dtype=torch.float32))
    if st.get("prompt_host") != tuple(prompt_ids):          # H2D of the prompt ids only when they change
        st["prompt"][:Lp].copy_(torch.tensor(prompt_ids, dtype=torch.int32), non_blocking=False)
        torch.cuda.current_stream(m.device).synchronize()
        st["prompt_host"] = tuple(prompt_ids)
    use_forced = forced_ids is not None
    if use_forced:
        st["forced"].copy_(forced_ids.to(device=m.device, dtype=torch.int32))
    lib = L.load()

    def enqueue():
        L.check(lib.vc_greedy_decode(C.byref(m.packed.gpt), pre.data_ptr(), n_seq, P, st["prompt"].data_ptr(), Lp, max_new, EOS,
                                     C.byref(st["cache"].c), st["ws"].data_ptr(), st["ws"].numel(), st["ids"].data_ptr(),
                                     st["lens"].data_ptr(), st["forced"].data_ptr() if use_forced else 0,
                                     st["logits"].data_ptr() if keep_logits else 0, L.current_stream(m.device)))

    gkey = "graph_forced" if use_forced else "graph"
    if not use_graph:
        enqueue()
    elif st[gkey] is None:
        enqueue()                              # eager warm-up (also sets func attributes outside capture)
        torch.cuda.current_stream(m.device).synchronize()
        g = torch.cuda.CUDAGraph()
        # kernel nodes keep the priority of the stream they were captured on (model.decode_priority: equal to the encoder's)
        cap = torch.cuda.Stream(device=m.device, priority=int(os.environ.get("VC_DECODE_PRIORITY", "0")))
        with torch.cuda.graph(g, stream=cap):
            enqueue()
        st[gkey] = g
        g.replay()
    else:
        st[gkey].replay()
    logits = st["logits"][:, :, : d["vocab"]] if keep_logits else None
    return st["ids"], st["lens"], logits


def hf_generate_ids(m, prefix: torch.Tensor, prompt_ids: List[int], *, max_new_tokens: int, num_beams: int = 1,
                    no_repeat_ngram_size: int = 3, repetition_penalty: float = 1.1, min_new_tokens: int = 8,
                    length_penalty: float = 1.0, do_sample: bool = False, temperature: float = 1.0, top_p: float = 1.0, top_k: int = 50,
                    generator=None):
    from .beam import beam_search_ids
    return beam_search_ids(m, prefix, prompt_ids, max_new_tokens=max_new_tokens, num_beams=num_beams,
                           no_repeat_ngram_size=no_repeat_ngram_size, repetition_penalty=repetition_penalty,
                           min_new_tokens=min_new_tokens, length_penalty=length_penalty, do_sample=do_sample,
                           temperature=temperature, top_p=top_p, top_k=top_k, generator=generator)
