"""Drives the UNMODIFIED reference modules (baseline/_ref, installed by baseline/install_reference.py) for bench.py's baseline arms.

  * `--impl reference` (CPU arm): `VideoCaptionModel` on the host cores, fp32, driven like
    core/scripts/benchmark_baseline.py:243-316 (`run_one_iteration`): encoder -> align_fn -> greedy KV-cache loop.  The loop
    of :160-240 is restated here only to drop its CUDA events / synchronisations, which do not exist on a CPU-only run.
  * `library_baseline` (GPU arm): the same modules moved to one B200 under bf16 autocast — cuBLAS / SDPA / eager PyTorch
    kernels, i.e. "the kernels the new code replaces"  — calling the reference's OWN
    `run_decoder_steps` (its CUDA events included) for the decode.

Offline shim: the GPU box has no network, no timm and no tokenizer files, so three names are patched at
run time before the reference modules are constructed: `GPT2LMHeadModel.from_pretrained` -> random-init `GPT2Config()`,
`GPT2TokenizerFast.from_pretrained` -> a stub with bos = eos = pad = 50256, and `video_encoder.vit_b_16` ->
torchvision `vit_b_16(weights=None)`.  No reference file is edited.  Weights are the synthetic state-dict both arms share.
"""
from __future__ import annotations

import os
import statistics
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "baseline" / "_ref"


class _StubTok:
    pad_token = None
    eos_token = "<|endoftext|>"
    bos_token_id = eos_token_id = pad_token_id = 50256

    def batch_decode(self, ids, skip_special_tokens=True):
        return ["" for _ in range(len(ids))]

    def decode(self, ids, skip_special_tokens=True):
        return ""


def available() -> str | None:
    """None if the reference tree is there, else the one-line reason."""
    if not (REF / "src" / "models" / "caption_model.py").exists():
        return f"{REF} is missing (run `python baseline/install_reference.py` in the build container)"
    return None


def build_model(state_dict: dict, prefix_len: int = 4):
    """The reference's VideoCaptionModel with the offline shim, loaded with `state_dict` (strict)."""
    import torch
    import torchvision
    import transformers
    from transformers import GPT2Config, GPT2LMHeadModel

    os.environ.setdefault("HF_HUB_OFFLINE", "1")
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    transformers.GPT2LMHeadModel.from_pretrained = classmethod(lambda cls, name, **kw: GPT2LMHeadModel(GPT2Config()))
    transformers.GPT2TokenizerFast.from_pretrained = classmethod(lambda cls, name, **kw: _StubTok())
    import src.models.video_encoder as ve
    ve.vit_b_16 = lambda weights=None: torchvision.models.vit_b_16(weights=None)
    from src.models.caption_model import VideoCaptionModel

    model = VideoCaptionModel(vit_enable_torch_compile=False, prefix_len=prefix_len).eval()
    own = model.state_dict()
    load = dict(state_dict)
    for k in own:                       # the unused torchvision classification head keeps its own init
        if k.startswith("encoder.backbone.model.heads."):
            load[k] = own[k]
    model.load_state_dict(load, strict=True)
    return model


def preprocess(frames_u8):
    """ToTensor + Normalize of core/preprocessing/frame_loader.py:34-40 on in-memory uint8 [B,T,H,W,3] frames (the Resize is the
    identity at 224 x 224; the JPEG decode of the harness has no counterpart for synthetic frames)."""
    import torch
    v = frames_u8.permute(0, 1, 4, 2, 3).to(torch.float32).div(255)
    mean = torch.tensor([0.485, 0.456, 0.406], dtype=torch.float32, device=v.device).view(1, 1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], dtype=torch.float32, device=v.device).view(1, 1, 3, 1, 1)
    return v.sub(mean).div(std)


def align(model, feat, ln_scale=0.6, in_weight=0.4):
    """align_fn of benchmark_baseline.py:267-283."""
    import torch
    emb = model.proj(feat)
    if emb.dim() == 2:
        emb = emb.unsqueeze(1)
    emb = torch.nn.functional.layer_norm(emb, emb.shape[-1:]) * ln_scale
    emb = emb * in_weight
    hidden = model.decoder.model.config.n_embd
    return model.decoder.mapper(emb).view(emb.size(0), model.decoder.prefix_len, hidden)


def greedy_cpu(model, prefix_embeds, max_new_tokens: int):
    """run_decoder_steps (benchmark_baseline.py:160-240) without its CUDA events; returns (token lists, per-step seconds)."""
    import torch
    gpt2 = model.decoder.model
    B = prefix_embeds.shape[0]
    prompt_ids = torch.tensor([[50256]], dtype=torch.long).expand(B, -1)
    full = torch.cat([prefix_embeds, gpt2.transformer.wte(prompt_ids)], dim=1)
    mask = torch.ones(full.shape[:2], dtype=torch.long)
    toks = [[] for _ in range(B)]
    past, nxt_in = None, full
    finished = torch.zeros(B, dtype=torch.bool)
    step_s = []
    for _ in range(max_new_tokens):
        t0 = time.perf_counter()
        out = gpt2(inputs_embeds=nxt_in, attention_mask=mask, past_key_values=past, use_cache=True, return_dict=True)
        step_s.append(time.perf_counter() - t0)
        nxt = torch.argmax(out.logits[:, -1, :], dim=-1)
        nxt = torch.where(finished, torch.full_like(nxt, 50256), nxt)
        for i, t in enumerate(nxt.tolist()):
            if not finished[i]:
                toks[i].append(t)
                if t == 50256:
                    finished[i] = True
        past = out.past_key_values
        if finished.all():
            break
        nxt_in = gpt2.transformer.wte(nxt).unsqueeze(1)
        mask = torch.cat([mask, torch.ones((B, 1), dtype=torch.long)], dim=1)
    return toks, step_s


def cpu_iteration(model, frames_u8, max_new_tokens: int):
    """One iteration of run_one_iteration (benchmark_baseline.py:243-316) on the host: returns seconds and stage split."""
    import torch
    with torch.inference_mode():
        t0 = time.perf_counter()
        video = preprocess(frames_u8)
        feat = model.encoder(video)
        t1 = time.perf_counter()
        prefix = align(model, feat)
        toks, step_s = greedy_cpu(model, prefix, max_new_tokens)
        t2 = time.perf_counter()
    return t2 - t0, (t1 - t0), step_s, toks


class _Profiler:
    """The attributes run_decoder_steps writes to (benchmark_baseline.py:206-208, :233-235)."""

    def __init__(self):
        self.token_step_ms, self.generated_lengths, self.last_texts = [], [], []


def gpu_library_baseline(state_dict, frames_u8_dev, max_new_tokens: int, iters: int = 3, warmup: int = 1):
    """The reference modules on one B200, eager PyTorch under bf16 autocast, one batch per iteration, driven like
    run_one_iteration; the decode is the reference's own run_decoder_steps (host sync + CUDA events per token)."""
    import torch
    model = build_model(state_dict).to(frames_u8_dev.device)
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    from core.scripts.benchmark_baseline import run_decoder_steps
    B = frames_u8_dev.shape[0]
    tot, enc, steps = [], [], []
    with torch.inference_mode(), torch.autocast("cuda", dtype=torch.bfloat16):
        for it in range(warmup + iters):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            video = preprocess(frames_u8_dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            feat = model.encoder(video)
            e1.record()
            prefix = align(model, feat)
            prof = _Profiler()
            run_decoder_steps(model, prefix, "", max_new_tokens, prof)
            torch.cuda.synchronize()
            if it >= warmup:
                tot.append(time.perf_counter() - t0)
                enc.append(e0.elapsed_time(e1))
                steps += prof.token_step_ms[1:]          # decode steps (token 0 is the prefill)
    steps.sort()
    del model
    torch.cuda.empty_cache()
    return {"captions_per_s": B / statistics.mean(tot), "ms_per_batch": statistics.mean(tot) * 1e3, "vit_encoder_ms": statistics.mean(enc),
            "decode_step_p50_us": steps[len(steps) // 2] * 1e3 if steps else None, "videos_per_batch": B, "iters": iters,
            "what": "unmodified reference modules (baseline/_ref) on this GPU: eager PyTorch, bf16 autocast, cuBLAS + SDPA kernels, "
                    "reference run_decoder_steps loop (host sync per token)"}
