"""Test infrastructure (never imported by the product): weights doctored so that the greedy loop's eos / finished branch
(core/scripts/benchmark_baseline.py:212-224) is exercised with decisive margins on random-init weights.

The tied embedding row of eos (50256) is set to beta * t * v and the row of token TRIGGER (7) to alpha * t * v, v a fixed random
unit vector, t the typical row norm, beta > alpha > 1.  Consequences, all with margins of several logits:
  * a row whose last hidden state has a positive component along v emits eos (some rows do at step 0, most never do);
  * feeding TRIGGER puts the next position's hidden state along v, so the NEXT argmax is eos: with teacher forcing
    (`forced_ids`) a row can be made to finish at any chosen step;
  * the prompt must not be bos (= eos): tests use prompt id PROMPT.
"""
from __future__ import annotations

import torch

EOS, TRIGGER, PROMPT = 50256, 7, 11


def doctor(sd: dict, dim: int, beta: float = 4.0, alpha: float = 3.0) -> dict:
    w = sd["decoder.model.transformer.wte.weight"]
    g = torch.Generator().manual_seed(99)
    v = torch.randn(dim, generator=g)
    v /= v.norm()
    t = w[:1000].norm(dim=1).mean()
    w[EOS] = beta * t * v
    w[TRIGGER] = alpha * t * v
    return sd


def prefixes(n: int, prefix_len: int, dim: int, seed: int = 5) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, prefix_len, dim, generator=g) * 0.3


def eos_margin(logits: torch.Tensor) -> torch.Tensor:
    """logit(eos) - best other logit, per row (positive = the row emits eos)."""
    other = logits.clone()
    other[..., EOS] = float("-inf")
    return logits[..., EOS] - other.max(dim=-1).values
