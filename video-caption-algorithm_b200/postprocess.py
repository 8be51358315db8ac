"""Host-side caption clean-up and candidate ranking: what `InferenceEngine.infer` applies to the decoded
strings of the three candidates (core/engine.py:75-83 -> core/postprocessing/text_cleaner.py:77-122 `clean_text`,
candidate_ranker.py:7-36 `score_sentence` / `select_best`).

Plain string work on a handful of short sentences per request — it stays on the host by design.  The behaviour is a rule table
(drop / strip / rewrite patterns and a sentence score), so it is written here as data plus a small interpreter; equality with the
reference's functions on a corpus of tricky strings is pinned by tests/golden/text_cleanup.json
(oracle/pin_text_against_reference.py) and checked in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import re
from typing import Iterable, Tuple

_I = re.IGNORECASE
FALLBACK = "Someone is in the scene."

# whole-string rejections: the caption is dropped (empty string)
_REJECT_FULL = [re.compile(r"[-_= \t]{6,}\.?"), re.compile(r'"\s*[^"]+\s*"\.?')]
_REJECT_LEAD = [re.compile(r"^\s*(https?://|www\.|<a\b|&lt;a\b)", _I), re.compile(r"^\s*(copyright\b)", _I),
                re.compile(r"^\s*(?:you are about to\b|click here\b|subscribe\b|available on youtube\b|watch live\b|find out\b|"
                           r"the video will\b|on the road\b)", _I)]
_REJECT_ANY = [re.compile(r"(</?\w+>|reddit\.com|pastebin|mailto:)", _I)]
_SPAM = re.compile(r"\b(click here|subscribe|report abuse|pastebin|official facebook|video will be)\b", _I)
_SPAM_TAIL = re.compile(r"\b(click here|subscribe|report abuse|pastebin|official facebook|video will be.*)$", _I)
# phrases removed outright, then prepositional chains normalised
_STRIP = [re.compile(p, _I) for p in (r"\bU\.S\.A?\.?\b", r"\bUSA\b", r"\bUnited States of America\b", r"\bUnited States\b", r"\bAmerica\b")]
_REWRITE = [(re.compile(r"\bin\s+the\s+front\s+of\b", _I), "in front of"), (re.compile(r"\bin\s+the\s+middle\s+of\b", _I), "in the middle of"),
            (re.compile(r"\bat\s+the\s+side\s+of\b", _I), "at the side of")]
_TAILS = [re.compile(r"\b(?:how|why|what|that|which)\b.*$", _I), re.compile(r"\bA\s+wonders\b.*$", _I)]
_SPACES = re.compile(r"\s{2,}")
_STUTTER = re.compile(r"\b(\w+)\b(?:\s+\1\b)+", _I)
_SENTENCE_SPLIT = re.compile(r"\s*(?<=\.|\!|\?)\s+")
# a token that marks the start of noise in a long caption: digits / slashes, dotted acronyms, codes like AB-12x, short all-caps words
_NOISE = [re.compile(r"[0-9/\\]"), re.compile(r"^(?:[A-Za-z]\.){2,}$"), re.compile(r"^[A-Z]{1,3}-[A-Za-z0-9]{1,6}$")]
_PUNCT = ",.;:!?()[]{}\"'`"


def score_sentence(text: str) -> float:
    """candidate_ranker.py:7-31: a Gaussian length prior around 12 tokens plus bonuses for a progressive verb, a copula and final
    punctuation, minus penalties for acronyms, spam phrases, very short sentences and the two stock fallbacks."""
    if not text:
        return -1e9
    n = len(text.split())
    s = -((n - 12.0) ** 2) / (2 * 4.0 * 4.0)
    s += 1.0 if re.search(r"\b\w+ing\b", text) else 0.0
    s += 0.5 if re.search(r"\b(?:is|are|was|were)\b", text) else 0.0
    s += 0.3 if text.endswith((".", "!", "?")) else 0.0
    s -= 1.5 if re.search(r"\b(?:[A-Z]\.){2,}\b", text) else 0.0
    s -= 1.5 if re.search(r"\b(click here|subscribe|report abuse|sign up|pastebin)\b", text, _I) else 0.0
    s -= 2.0 if n < 4 else 0.0
    s -= 0.8 if text.strip().lower() in {"someone is sitting.", "someone is in the scene."} else 0.0
    return s


def select_best(candidates: Iterable[Tuple[str, str]]):
    """candidate_ranker.py:34-36: (key, text, score) of the best-scoring candidate; the first one wins a tie."""
    best = None
    for key, value in candidates:
        sc = score_sentence(value)
        if best is None or sc > best[2]:
            best = (key, value, sc)
    if best is None:
        raise IndexError("select_best: no candidates")
    return best


def _cut_at_noise(text: str) -> str:
    words = text.split()
    keep = len(words)
    for i, w in enumerate(words):
        core = w.strip(_PUNCT)
        if core and (any(p.search(core) for p in _NOISE) or (len(core) <= 3 and core.isupper())):
            keep = i
            break
    out = " ".join(words[:keep]).strip()
    return out + "." if out and out[-1] not in ".!?" else out


def _sitting_needs_a_place(text: str) -> str:
    low = text.strip().lower()
    if re.match(r"^someone\s+is\b", low):          # (the reference returns here for every "someone is ..." sentence)
        return text
    if re.match(r"^someone\s+is\s+sitting\s*\.?$", low):
        return "Someone is sitting on a chair."
    if re.match(r"^someone\s+is\s+sitting\b", low) and not re.search(r"\b(in|on|at|by|with|near)\b", low):
        return text.rstrip(". ") + " on a chair."
    return text


def clean_text(raw: str) -> str:
    """text_cleaner.py:77-122: raw decoder output -> one subtitle-like sentence ('' when the caption is boilerplate)."""
    text = (raw or "").strip()
    if _REJECT_FULL[0].fullmatch(text):
        return ""
    text = re.sub(r"^\s*[-_= \t]{2,}\s*", "", text)
    if any(p.match(text) for p in _REJECT_LEAD[:2]) or _REJECT_FULL[1].fullmatch(text):
        return ""
    if _REJECT_LEAD[2].match(text) or any(p.search(text) for p in _REJECT_ANY):
        return ""
    spam = bool(_SPAM.search(text))
    text = _SPAM_TAIL.sub("", text).strip()
    for p in _STRIP:
        text = p.sub("", text)
    text = _SPACES.sub(" ", text).strip()
    for p, repl in _REWRITE:
        text = p.sub(repl, text)
    text = _SPACES.sub(" ", text)
    if len(text.split()) >= 10:
        text = _cut_at_noise(text)
    for p in _TAILS:
        text = p.sub("", text).strip()
    text = text or FALLBACK
    if spam and len(text.split()) <= 2:
        text = FALLBACK
    text = _sitting_needs_a_place(text)
    text = _STUTTER.sub(r"\1", text)
    text = _SPACES.sub(" ", text).strip()
    if text and text[0].isalpha():
        text = text[0].upper() + text[1:]
    if text and text[-1] not in ".!?":
        text += "."
    parts = [c.strip() for c in _SENTENCE_SPLIT.split(text) if c.strip()]
    return parts[0] if parts and parts[0] else text
