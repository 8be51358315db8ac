"""CPU oracle for the video-caption hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

A plain fp32 restatement (torch CPU tensor ops, no nn.Module from torchvision /
transformers / timm) of what the reference computes on the path
preprocess -> ViT encode -> temporal pool + projection -> prefix -> GPT-2
greedy / beam decode.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import this module; the
product package never does (tests/test_layout.py enforces it).

Pinning: the reference has no tests or golden vectors for this path
(SURVEY.md §4, §8c).  The oracle is pinned against the reference's OWN modules
(src/models/*.py driven exactly like core/scripts/benchmark_baseline.py) run in
the build container by `oracle/pin_against_reference.py`; the outputs are
committed under `tests/golden/` and `tests/test_oracle_golden.py` checks this
file against them on every CPU run.

Every function cites the reference lines it restates (paths relative to the
reference repo root); arithmetic that lives in third-party code is cited by
package and symbol (torchvision 0.20.1 / transformers 4.57.1 pinned in the
reference's requirements.txt:36-38; 0.26.0 / 5.5.0 installed here).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


# --------------------------------------------------------------------------
# a1  preprocessing — core/preprocessing/frame_loader.py:19-49
# --------------------------------------------------------------------------
def sample_frame_indices(n_files: int, num_frames: int) -> list[int]:
    """frame_loader.py:31-32: step=max(n//T,1); files[::step][:T] (no padding)."""
    step = max(n_files // num_frames, 1)
    return list(range(0, n_files, step))[:num_frames]


def normalize_lut() -> torch.Tensor:
    """[3,256] fp32 table of ToTensor+Normalize applied to every uint8 value.

    frame_loader.py:34-40: ToTensor is `u8.to(float32).div(255)`
    (torchvision.transforms.functional.to_tensor), Normalize is
    `.sub(mean).div(std)` per channel (functional.normalize) — sequentially
    rounded fp32 ops, restated here with the same torch ops.
    """
    v = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(3, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(3, 1)
    return v.view(1, 256).sub(mean).div(std).contiguous()


def pillow_bilinear_coeffs(in_size: int, out_size: int):
    """Pillow `precompute_coeffs` + `normalize_coeffs_8bpc` for the bilinear filter (src/libImaging/Resample.c; Pillow 10.0.1
    pinned in the reference's requirements.txt:70, 12.2.0 installed here — the routine is unchanged between them).
    Support = max(in/out, 1) (antialiasing when shrinking); per output index a window [xmin, xmin+n) and weights
    normalised in double precision, then rounded to 22 fractional bits.  -> (int weights [out][ksize], (xmin, n) [out])."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    weights, bounds = [], []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        xmin = 0 if xmin < 0 else xmin
        xmax = int(center + support + 0.5)
        xmax = in_size if xmax > in_size else xmax
        n = xmax - xmin
        k, ww = [], 0.0
        for x in range(n):
            v = (x + xmin - center + 0.5) * (1.0 / filterscale)
            v = -v if v < 0.0 else v
            w = 1.0 - v if v < 1.0 else 0.0
            k.append(w)
            ww += w
        if ww != 0.0:
            k = [w / ww for w in k]
        k += [0.0] * (ksize - n)
        weights.append([int(-0.5 + w * (1 << 22)) if w < 0 else int(0.5 + w * (1 << 22)) for w in k])
        bounds.append((xmin, n))
    return weights, bounds


def resize_bilinear_u8(frames_hwc: torch.Tensor, out_h: int, out_w: int) -> torch.Tensor:
    """frame_loader.py:36 `transforms.Resize((image_size, image_size))` on a PIL RGB image = PIL.Image.resize(BILINEAR) =
    Pillow ImagingResample (8 bpc): horizontal pass, uint8 intermediate, vertical pass; each output byte is
    clip8((2^21 + sum in*w) >> 22) in integer arithmetic.  uint8 [...,H,W,3] -> uint8 [...,out_h,out_w,3]."""
    x = frames_hwc.to(torch.int64)
    H, W = x.shape[-3], x.shape[-2]
    if W != out_w:
        wts, bnd = pillow_bilinear_coeffs(W, out_w)
        cols = []
        for xo in range(out_w):
            xmin, n = bnd[xo]
            k = torch.tensor(wts[xo][:n], dtype=torch.int64).view(n, 1)
            acc = (x[..., xmin:xmin + n, :] * k).sum(dim=-2) + (1 << 21)
            cols.append((acc >> 22).clamp_(0, 255))
        x = torch.stack(cols, dim=-2)
    if H != out_h:
        wts, bnd = pillow_bilinear_coeffs(H, out_h)
        rows = []
        for yo in range(out_h):
            ymin, n = bnd[yo]
            k = torch.tensor(wts[yo][:n], dtype=torch.int64).view(n, 1, 1)
            acc = (x[..., ymin:ymin + n, :, :] * k).sum(dim=-3) + (1 << 21)
            rows.append((acc >> 22).clamp_(0, 255))
        x = torch.stack(rows, dim=-3)
    return x.to(torch.uint8)


def preprocess_u8(frames_hwc: torch.Tensor) -> torch.Tensor:
    """uint8 [...,H,W,3] (already image_size x image_size, so Resize is the
    identity: frame_loader.py:36, PIL returns a copy) -> fp32 [...,3,H,W]."""
    x = frames_hwc.to(torch.float32).div(255)
    x = x.movedim(-1, -3)
    mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32).view(3, 1, 1)
    std = torch.tensor(IMAGENET_STD, dtype=torch.float32).view(3, 1, 1)
    return x.sub(mean).div(std).contiguous()


# --------------------------------------------------------------------------
# a2  ViT frame encoder — src/models/video_encoder.py:288-326
# --------------------------------------------------------------------------
def _vit_keys(sd: dict) -> dict:
    """Resolve the two key layouts the reference can produce."""
    if "encoder.backbone.model.class_token" in sd:  # torchvision fallback, video_encoder.py:84-103
        p = "encoder.backbone.model."
        n = sum(1 for k in sd if k.startswith(p + "encoder.layers.") and k.endswith("ln_1.weight"))
        blk = lambda i: f"{p}encoder.layers.encoder_layer_{i}."
        return dict(cls=p + "class_token", pos=p + "encoder.pos_embedding", pw=p + "conv_proj.weight",
                    pb=p + "conv_proj.bias", n=n, blk=blk, ln1="ln_1", qkv_w="self_attention.in_proj_weight",
                    qkv_b="self_attention.in_proj_bias", out="self_attention.out_proj", ln2="ln_2",
                    fc1="mlp.0", fc2="mlp.3", lnf=p + "encoder.ln", gelu="erf")
    p = "encoder.backbone."  # timm, video_encoder.py:69-80
    n = sum(1 for k in sd if k.startswith(p + "blocks.") and k.endswith("norm1.weight"))
    blk = lambda i: f"{p}blocks.{i}."
    return dict(cls=p + "cls_token", pos=p + "pos_embed", pw=p + "patch_embed.proj.weight",
                pb=p + "patch_embed.proj.bias", n=n, blk=blk, ln1="norm1", qkv_w="attn.qkv.weight",
                qkv_b="attn.qkv.bias", out="attn.proj", ln2="norm2", fc1="mlp.fc1", fc2="mlp.fc2",
                lnf=p + "norm", gelu="tanh")


def vit_tokens(sd: dict, x: torch.Tensor, heads: int, gelu: Optional[str] = None,
               taps: Optional[dict] = None) -> torch.Tensor:
    """fp32 [n,3,H,W] -> final-LayerNorm tokens [n,N,D].

    torchvision path (the one runnable in the build image): `_process_input`
    (conv k=s=patch, reshape, permute), class token concat
    (video_encoder.py:90-95), then torchvision `Encoder.forward`
    (+pos_embedding, EncoderBlock x L, `ln`) — vision_transformer.py
    Encoder/EncoderBlock/MLPBlock: x = in + MHA(ln_1(in)); out = x + MLP(ln_2(x)),
    LayerNorm eps 1e-6, exact-erf GELU.  timm path: same graph, tanh GELU after
    the reference's patch at video_encoder.py:123-134.
    """
    k = _vit_keys(sd)
    gelu = gelu or k["gelu"]
    w = sd[k["pw"]]
    D, patch = w.shape[0], w.shape[-1]
    n = x.shape[0]
    t = F.conv2d(x, w, sd[k["pb"]], stride=patch)           # [n,D,g,g]
    t = t.reshape(n, D, -1).permute(0, 2, 1)                 # [n,g*g,D]
    t = torch.cat([sd[k["cls"]].expand(n, -1, -1), t], 1)    # [n,N,D]
    t = t + sd[k["pos"]]
    hd = D // heads
    for i in range(k["n"]):
        b = k["blk"](i)
        h = F.layer_norm(t, (D,), sd[b + k["ln1"] + ".weight"], sd[b + k["ln1"] + ".bias"], 1e-6)
        qkv = F.linear(h, sd[b + k["qkv_w"]], sd[b + k["qkv_b"]])
        q, kk, v = qkv.view(n, -1, 3, heads, hd).permute(2, 0, 3, 1, 4)  # each [n,heads,N,hd]
        att = torch.softmax((q @ kk.transpose(-1, -2)) * (hd ** -0.5), dim=-1)
        o = (att @ v).transpose(1, 2).reshape(n, -1, D)
        t = t + F.linear(o, sd[b + k["out"] + ".weight"], sd[b + k["out"] + ".bias"])
        h = F.layer_norm(t, (D,), sd[b + k["ln2"] + ".weight"], sd[b + k["ln2"] + ".bias"], 1e-6)
        h = F.linear(h, sd[b + k["fc1"] + ".weight"], sd[b + k["fc1"] + ".bias"])
        h = F.gelu(h, approximate="tanh" if gelu == "tanh" else "none")
        t = t + F.linear(h, sd[b + k["fc2"] + ".weight"], sd[b + k["fc2"] + ".bias"])
        if taps is not None:
            taps[f"block{i}"] = t
    return F.layer_norm(t, (D,), sd[k["lnf"] + ".weight"], sd[k["lnf"] + ".bias"], 1e-6)


def encode(sd: dict, video: torch.Tensor, heads: int = 12, gelu: Optional[str] = None) -> torch.Tensor:
    """`ViTFrameEncoder.forward`: [B,T,3,H,W] fp32 -> [B,video_dim] fp32.

    video_encoder.py:293-294 fold T into batch; :314 -> :256-258 cls token,
    temporal mean; :316 proj Linear; l2norm off (caption_model.py:46); :323 float.
    """
    B, T = video.shape[:2]
    tok = vit_tokens(sd, video.reshape(B * T, *video.shape[2:]), heads, gelu)
    pooled = tok.reshape(B, T, tok.shape[1], tok.shape[2])[:, :, 0, :].mean(dim=1)
    return F.linear(pooled, sd["encoder.proj.weight"], sd["encoder.proj.bias"]).float()


# --------------------------------------------------------------------------
# a4-a6  Cross_Modal_Alignment — core/engine.py:44-50, text_decoder.py:36-45,69
# --------------------------------------------------------------------------
def visual_prefix(sd: dict, feat: torch.Tensor, ln_scale: float = 0.6, in_weight: float = 0.4,
                  prefix_len: int = 4) -> torch.Tensor:
    """[B,video_dim] -> [B,P,H]: proj=Identity (caption_model.py:67), unsqueeze,
    F.layer_norm(no affine, eps 1e-5)*ln_scale, *in_weight (engine.py:45-50 /
    benchmark_baseline.py:267-274), mapper Linear + eval Dropout, view
    (text_decoder.py:69 / benchmark_baseline.py:276-278)."""
    emb = feat.unsqueeze(1)
    if ln_scale is not None and ln_scale > 0:
        emb = F.layer_norm(emb, emb.shape[-1:]) * ln_scale
    if in_weight is not None and in_weight > 0:
        emb = emb * in_weight
    m = F.linear(emb, sd["decoder.mapper.0.weight"], sd["decoder.mapper.0.bias"])
    return m.view(feat.shape[0], prefix_len, -1)


# --------------------------------------------------------------------------
# a9  GPT-2 forward with KV cache — transformers modeling_gpt2.GPT2LMHeadModel
# --------------------------------------------------------------------------
def _gelu_new(x: torch.Tensor) -> torch.Tensor:
    """transformers.activations.NewGELUActivation."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def gpt2_forward(sd: dict, inputs_embeds: torch.Tensor, past: Optional[list], heads: int = 12,
                 last_only: bool = True):
    """inputs_embeds [B,L,H], past = list of (K,V) [B,heads,S,hd] or None ->
    (logits [B,(1|L),V] fp32, new past).

    GPT2Model.forward: h = inputs_embeds + wpe(arange(L)+past_len) (all-ones
    mask => plain positions); GPT2Block: x += c_proj(attn(ln_1 x));
    x += mlp.c_proj(gelu_new(c_fc(ln_2 x))); LayerNorm eps 1e-5; Conv1D is
    x @ W[in,out] + b; attention scale hd**-0.5, causal; ln_f; lm_head = wte^T.
    HF computes logits for all L positions (benchmark_baseline.py:197-203 then
    takes [:, -1, :]); `last_only` skips the unused rows.
    """
    g = "decoder.model.transformer."
    B, L, H = inputs_embeds.shape
    n_layer = sum(1 for k in sd if k.startswith(g + "h.") and k.endswith("ln_1.weight"))
    past_len = 0 if past is None else past[0][0].shape[2]
    pos = torch.arange(past_len, past_len + L)
    h = inputs_embeds + sd[g + "wpe.weight"][pos]
    hd = H // heads
    new_past = []
    S = past_len + L
    causal = torch.ones(L, S, dtype=torch.bool).tril(diagonal=past_len)
    for i in range(n_layer):
        p = f"{g}h.{i}."
        x = F.layer_norm(h, (H,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = x @ sd[p + "attn.c_attn.weight"] + sd[p + "attn.c_attn.bias"]
        q, k, v = qkv.split(H, dim=2)
        q = q.view(B, L, heads, hd).transpose(1, 2)
        k = k.view(B, L, heads, hd).transpose(1, 2)
        v = v.view(B, L, heads, hd).transpose(1, 2)
        if past is not None:
            k = torch.cat([past[i][0], k], dim=2)
            v = torch.cat([past[i][1], v], dim=2)
        new_past.append((k, v))
        s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
        s = s.masked_fill(~causal, float("-inf"))
        o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, L, H)
        h = h + (o @ sd[p + "attn.c_proj.weight"] + sd[p + "attn.c_proj.bias"])
        x = F.layer_norm(h, (H,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        x = _gelu_new(x @ sd[p + "mlp.c_fc.weight"] + sd[p + "mlp.c_fc.bias"])
        h = h + (x @ sd[p + "mlp.c_proj.weight"] + sd[p + "mlp.c_proj.bias"])
    if last_only:
        h = h[:, -1:, :]
    h = F.layer_norm(h, (H,), sd[g + "ln_f.weight"], sd[g + "ln_f.bias"], 1e-5)
    return h @ sd[g + "wte.weight"].t(), new_past


def build_inputs(sd: dict, prefix: torch.Tensor, prompt_ids: torch.Tensor) -> torch.Tensor:
    """text_decoder.py:60-74 / benchmark_baseline.py:176-180: wte(prompt) expanded
    to the batch, cat([prefix, tok_emb], dim=1)."""
    tok = sd["decoder.model.transformer.wte.weight"][prompt_ids]
    if tok.shape[0] == 1 and prefix.shape[0] > 1:
        tok = tok.expand(prefix.shape[0], -1, -1)
    return torch.cat([prefix, tok], dim=1)


# --------------------------------------------------------------------------
# a8  benchmark greedy loop — core/scripts/benchmark_baseline.py:160-240
# --------------------------------------------------------------------------
def greedy_decode(sd: dict, prefix: torch.Tensor, prompt_ids: torch.Tensor, max_new_tokens: int,
                  eos: int = 50256, heads: int = 12, forced_ids: Optional[torch.Tensor] = None,
                  keep_logits: bool = False):
    """Returns (ids [B,max_new] int64 padded with eos, lengths [B], logits list).

    Step 0 is the prefill over P+Lp embeds; argmax of logits[:, -1] (ties ->
    lowest index, torch.argmax); finished rows are forced to eos (:212-214);
    a row's tokens are appended until and including its first eos (:216-221);
    the loop stops when every row has finished (:224).  No repetition penalty /
    n-gram block here.  `forced_ids` (teacher forcing) replaces the fed-back
    token, for tolerance checks that must not diverge.
    """
    B = prefix.shape[0]
    x = build_inputs(sd, prefix, prompt_ids)
    past = None
    finished = torch.zeros(B, dtype=torch.bool)
    ids = torch.full((B, max_new_tokens), eos, dtype=torch.int64)
    lengths = torch.zeros(B, dtype=torch.int64)
    all_logits = []
    wte = sd["decoder.model.transformer.wte.weight"]
    for step in range(max_new_tokens):
        logits, past = gpt2_forward(sd, x, past, heads)
        logits = logits[:, -1, :]
        if keep_logits:
            all_logits.append(logits.clone())
        nxt = torch.argmax(logits, dim=-1)
        nxt = torch.where(finished, torch.full_like(nxt, eos), nxt)
        live = ~finished
        ids[live, step] = nxt[live]
        lengths += live.long()
        finished = finished | (nxt == eos)
        if bool(finished.all()) and forced_ids is None:
            break
        feed = nxt if forced_ids is None else forced_ids[:, step]
        x = wte[feed].unsqueeze(1)
    return ids, lengths, all_logits


# --------------------------------------------------------------------------
# a10  HF generate: logits processors + beam search
#      transformers/generation/logits_process.py, generation/utils.py `_beam_search`
# --------------------------------------------------------------------------
def apply_processors(scores: torch.Tensor, seqs: torch.Tensor, cur_len: int, *, repetition_penalty: float,
                     no_repeat_ngram_size: int, min_new_tokens: int, eos: int) -> torch.Tensor:
    """scores [n,V] fp32 (log-probs for beams, raw logits for greedy), seqs [n,cur_len]
    generated so far.  Order as HF builds it (generation/utils.py
    `_get_logits_processor`): RepetitionPenalty -> NoRepeatNGram -> MinNewTokens.
    """
    scores = scores.clone()
    n = scores.shape[0]
    if repetition_penalty != 1.0 and cur_len > 0:
        # RepetitionPenaltyLogitsProcessor: score<0 ? score*p : score/p on seen ids
        sc = torch.gather(scores, 1, seqs)
        sc = torch.where(sc < 0, sc * repetition_penalty, sc / repetition_penalty)
        scores.scatter_(1, seqs, sc)
    if no_repeat_ngram_size > 0 and cur_len + 1 >= no_repeat_ngram_size:
        # NoRepeatNGramLogitsProcessor (_calc_banned_ngram_tokens)
        k = no_repeat_ngram_size
        for r in range(n):
            row = seqs[r].tolist()
            tail = tuple(row[cur_len + 1 - k:cur_len]) if k > 1 else ()
            for s in range(cur_len - k + 1):
                if tuple(row[s:s + k - 1]) == tail:
                    scores[r, row[s + k - 1]] = float("-inf")
    if min_new_tokens > 0 and cur_len < min_new_tokens:
        # MinNewTokensLengthLogitsProcessor (prompt length 0: prompt went in as embeds)
        scores[:, eos] = float("-inf")
    return scores


def sample_warpers(scores: torch.Tensor, temperature: float, top_p: float, top_k: int = 50) -> torch.Tensor:
    """do_sample=True branch of text_decoder.py:131-144: transformers TemperatureLogitsWarper (`scores / temperature`), then
    TopKLogitsWarper — the reference passes no top_k, so GenerationConfig's default top_k = 50 is in force
    (generation/utils.py `_get_logits_processor`: temperature, top_k, top_p in that order; logits_process.py: remove everything
    below the k-th largest score) — then TopPLogitsWarper (sort ascending, cumulative softmax, remove while cumsum <= 1 - top_p,
    keep at least the last = most probable token, scatter the mask back, fill with -inf).  The draw itself
    (`torch.multinomial`) uses the framework RNG and is not part of the oracle.  Pinned against the installed transformers
    classes by tests/test_oracle_golden.py::test_sample_warpers_equal_transformers_processors."""
    scores = scores / temperature
    if top_k and top_k > 0:
        k = min(int(top_k), scores.size(-1))
        indices_to_remove = scores < torch.topk(scores, k)[0][..., -1, None]
        scores = scores.masked_fill(indices_to_remove, float("-inf"))
    if top_p < 1.0:
        sorted_logits, sorted_indices = torch.sort(scores, descending=False)
        cumulative_probs = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
        sorted_indices_to_remove = cumulative_probs <= (1 - top_p)
        sorted_indices_to_remove[..., -1:] = 0
        indices_to_remove = sorted_indices_to_remove.scatter(1, sorted_indices, sorted_indices_to_remove)
        scores = scores.masked_fill(indices_to_remove, float("-inf"))
    return scores


def beam_search(sd: dict, prefix: torch.Tensor, prompt_ids: torch.Tensor, *, num_beams: int, max_new_tokens: int,
                no_repeat_ngram_size: int = 3, repetition_penalty: float = 1.1, min_new_tokens: int = 8,
                length_penalty: float = 1.0, eos: int = 50256, heads: int = 12, step_logits_fn=None):
    """HF `GenerationMixin._beam_search` with the kwargs of text_decoder.py:131-144
    (early_stopping=False, length_penalty=1.0, pad=eos) restated from the installed
    transformers source.  Returns (ids [B,max_new] padded with eos,
    lengths [B]) of the best finished beam per video.

    `step_logits_fn(step, beam_idx_or_None, tokens_or_None) -> logits [B*nb,V]`
    lets a test drive the same bookkeeping from another model's logits.
    """
    B = prefix.shape[0]
    nb = num_beams
    V = sd["decoder.model.transformer.wte.weight"].shape[0]
    wte = sd["decoder.model.transformer.wte.weight"]
    state = {"past": None}

    def own_logits(step, beam_idx, tokens):
        if step == 0:
            x = build_inputs(sd, prefix.repeat_interleave(nb, dim=0), prompt_ids)
        else:
            state["past"] = [(k[beam_idx], v[beam_idx]) for k, v in state["past"]]
            x = wte[tokens].unsqueeze(1)
        lg, state["past"] = gpt2_forward(sd, x, state["past"], heads)
        return lg[:, -1, :]

    fn = step_logits_fn or own_logits
    NEG = -1.0e9
    running_scores = torch.zeros(B, nb)
    running_scores[:, 1:] = NEG
    running_seqs = torch.full((B, nb, max_new_tokens), eos, dtype=torch.int64)
    fin_seqs = running_seqs.clone()
    fin_scores = torch.full((B, nb), NEG)
    fin_done = torch.zeros(B, nb, dtype=torch.bool)
    fin_len = torch.zeros(B, nb, dtype=torch.int64)
    unsatisfied = torch.ones(B, 1, dtype=torch.bool)          # is_early_stop_heuristic_unsatisfied (sticky)
    beam_idx_flat, tokens_flat = None, None
    cur_len = 0
    keep = 2 * nb                                              # (n_eos + 1) * num_beams
    in_top = torch.arange(keep).view(1, keep) < nb             # top_num_beam_mask
    while True:
        logits = fn(cur_len, beam_idx_flat, tokens_flat).float()
        logp = torch.log_softmax(logits, dim=-1)
        flat = running_seqs.view(B * nb, max_new_tokens)[:, :cur_len]
        logp = apply_processors(logp, flat, cur_len, repetition_penalty=repetition_penalty,
                                no_repeat_ngram_size=no_repeat_ngram_size, min_new_tokens=min_new_tokens, eos=eos)
        cand = (logp.view(B, nb, V) + running_scores.unsqueeze(-1)).view(B, nb * V)
        top_scores, top_idx = torch.topk(cand, keep, dim=1)    # _get_top_k_continuations
        top_beam = top_idx // V
        top_tok = top_idx % V
        cand_seqs = torch.gather(running_seqs, 1, top_beam.unsqueeze(-1).expand(-1, -1, max_new_tokens)).clone()
        cand_seqs[:, :, cur_len] = top_tok
        new_len = cur_len + 1
        hit_stop = (top_tok == eos) | (new_len >= max_new_tokens)   # EosTokenCriteria | MaxLengthCriteria
        # _get_running_beams_for_next_iteration
        run_rank = top_scores + hit_stop.float() * NEG
        running_scores, pick = torch.topk(run_rank, nb, dim=1)
        running_seqs = torch.gather(cand_seqs, 1, pick.unsqueeze(-1).expand(-1, -1, max_new_tokens))
        running_beam = torch.gather(top_beam, 1, pick)
        running_tok = torch.gather(top_tok, 1, pick)
        # _update_finished_beams
        newly = hit_stop & in_top
        f_scores = top_scores / (float(new_len) ** length_penalty)
        f_scores = f_scores + (~unsatisfied).float() * NEG
        f_scores = f_scores + (~newly).float() * NEG
        merged_scores = torch.cat([fin_scores, f_scores], dim=1)
        merged_seqs = torch.cat([fin_seqs, cand_seqs], dim=1)
        merged_done = torch.cat([fin_done, newly], dim=1)
        merged_len = torch.cat([fin_len, torch.full((B, keep), new_len, dtype=torch.int64)], dim=1)
        fin_scores, sel = torch.topk(merged_scores, nb, dim=1)
        fin_seqs = torch.gather(merged_seqs, 1, sel.unsqueeze(-1).expand(-1, -1, max_new_tokens))
        fin_done = torch.gather(merged_done, 1, sel)
        fin_len = torch.gather(merged_len, 1, sel)
        cur_len = new_len
        beam_idx_flat = (running_beam + torch.arange(B).view(B, 1) * nb).view(-1)
        tokens_flat = running_tok.view(-1)
        # _check_early_stop_heuristic (early_stopping=False)
        best_running = running_scores[:, :1] / (float(cur_len) ** length_penalty)
        worst_fin = torch.where(fin_done, fin_scores.min(dim=1, keepdim=True).values, torch.full_like(fin_scores, NEG))
        unsatisfied = unsatisfied & (best_running > worst_fin).any(dim=-1, keepdim=True)
        # _beam_search_has_unfinished_sequences
        if not (bool(unsatisfied.any()) and not bool(hit_stop.all())):
            break
    ids = torch.full((B, max_new_tokens), eos, dtype=torch.int64)
    out_len = fin_len[:, 0].clone()
    for b in range(B):
        n_tok = int(out_len[b])
        ids[b, :n_tok] = fin_seqs[b, 0, :n_tok]
    return ids, out_len


# --------------------------------------------------------------------------
# whole path (a11 minus host string work): frames -> ids
# --------------------------------------------------------------------------
def caption_ids(sd: dict, frames_u8: torch.Tensor, *, vit_heads: int = 12, gpt_heads: int = 12,
                prompt_ids: Optional[torch.Tensor] = None, max_new_tokens: int = 20, num_beams: int = 1,
                ln_scale: float = 0.6, in_weight: float = 0.4, prefix_len: int = 4, eos: int = 50256):
    """uint8 [B,T,H,W,3] -> (ids, lengths, feat, prefix) with the benchmark's greedy
    semantics (num_beams=1) or HF beam search (num_beams>1)."""
    if prompt_ids is None:
        prompt_ids = torch.tensor([[eos]], dtype=torch.int64)  # text_decoder.py:122
    video = preprocess_u8(frames_u8)
    feat = encode(sd, video, vit_heads)
    prefix = visual_prefix(sd, feat, ln_scale, in_weight, prefix_len)
    if num_beams == 1:
        ids, lengths, _ = greedy_decode(sd, prefix, prompt_ids, max_new_tokens, eos, gpt_heads)
    else:
        ids, lengths = beam_search(sd, prefix, prompt_ids, num_beams=num_beams, max_new_tokens=max_new_tokens,
                                   eos=eos, heads=gpt_heads)
    return ids, lengths, feat, prefix
