"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference; the GPU box has no
copy):   python oracle/pin_against_reference.py

What it does:
  * imports the reference's `src.models.caption_model.VideoCaptionModel` from
    /root/reference with three offline monkey-patches (HF `from_pretrained` ->
    random-init config, a stub tokenizer, torchvision `vit_b_16(weights=None)`)
    — the reference source itself is untouched and not copied;
  * loads the synthetic state-dict of `video-caption-algorithm_b200/synthetic.py`
    (strict) so the fixture can be regenerated from a seed anywhere;
  * drives the model exactly like core/scripts/benchmark_baseline.py:265-289
    (encoder, align_fn, the greedy KV-cache loop :160-240 minus CUDA events) and
    like core/engine.py:52-61 (`decoder.generate` with beams);
  * also runs the reference under bf16 autocast to record the reference's own
    bf16-vs-fp32 error, which is what the GPU tolerances are anchored to;
  * writes small outputs only (features, prefix, ids, sub-sampled logits).
"""
from __future__ import annotations

import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
REF = Path(os.environ.get("VC_REFERENCE", "/root/reference"))
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REF))
os.environ.setdefault("HF_HUB_OFFLINE", "1")

import vcb200  # noqa: E402
from vcb200 import synthetic  # noqa: E402

LOGIT_STRIDE = 97  # sub-sampling of the 50257-wide logits kept in the fixture


class _StubTok:
    """Stands in for GPT2TokenizerFast (no BPE files offline)."""
    pad_token = None
    eos_token = "<|endoftext|>"
    bos_token_id = eos_token_id = pad_token_id = 50256
    captured = None

    def batch_decode(self, ids, skip_special_tokens=True):
        _StubTok.captured = ids.clone()
        return ["" for _ in range(ids.shape[0])]

    def decode(self, ids, skip_special_tokens=True):
        return ""


def build_reference_model(arch_name: str, seed: int):
    import transformers
    import torchvision
    from transformers import GPT2Config, GPT2LMHeadModel

    a = synthetic.ARCHS[arch_name]
    cfg = GPT2Config(n_embd=a.gpt_dim, n_layer=a.gpt_layers, n_head=a.gpt_heads)
    transformers.GPT2LMHeadModel.from_pretrained = classmethod(lambda cls, name, **kw: GPT2LMHeadModel(cfg))
    transformers.GPT2TokenizerFast.from_pretrained = classmethod(lambda cls, name, **kw: _StubTok())
    import src.models.video_encoder as ve

    def _vit(weights=None):
        m = torchvision.models.vit_b_16(weights=None)
        if a.vit_layers != 12:  # "tiny" cut: keep the first L blocks of the same module
            layers = m.encoder.layers
            for i in range(a.vit_layers, 12):
                delattr(layers, f"encoder_layer_{i}")
        return m

    ve.vit_b_16 = _vit
    from src.models.caption_model import VideoCaptionModel

    model = VideoCaptionModel(vit_enable_torch_compile=False, prefix_len=a.prefix_len).eval()
    sd = synthetic.make_state_dict(a, seed=seed, layout="torchvision")
    own = model.state_dict()
    load = {k: v for k, v in sd.items()}
    # the unused torchvision classification head stays at its own init
    for k in own:
        if k.startswith("encoder.backbone.model.heads."):
            load[k] = own[k]
    model.load_state_dict(load, strict=True)
    return model, sd


@torch.inference_mode()
def reference_greedy(model, prefix, max_new_tokens):
    """core/scripts/benchmark_baseline.py:160-240 with the timing calls removed."""
    gpt2 = model.decoder.model
    B = prefix.shape[0]
    prompt_ids = torch.tensor([[50256]], dtype=torch.long).expand(B, -1)
    full = torch.cat([prefix, gpt2.transformer.wte(prompt_ids)], dim=1)
    mask = torch.ones(full.shape[:2], dtype=torch.long)
    toks = [[] for _ in range(B)]
    past, nxt_in = None, full
    finished = torch.zeros(B, dtype=torch.bool)
    logits_all = []
    for _ in range(max_new_tokens):
        out = gpt2(inputs_embeds=nxt_in, attention_mask=mask, past_key_values=past, use_cache=True, return_dict=True)
        logits = out.logits[:, -1, :]
        logits_all.append(logits.float().clone())
        nxt = torch.argmax(logits, dim=-1)
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
    ids = torch.full((B, max_new_tokens), 50256, dtype=torch.int64)
    for i, t in enumerate(toks):
        ids[i, :len(t)] = torch.tensor(t)
    return ids, torch.tensor([len(t) for t in toks]), logits_all


def align(model, feat, ln_scale=0.6, in_weight=0.4):
    """benchmark_baseline.py:267-278."""
    emb = model.proj(feat).unsqueeze(1)
    emb = torch.nn.functional.layer_norm(emb, emb.shape[-1:]) * ln_scale
    emb = emb * in_weight
    hid = model.decoder.model.config.n_embd
    return emb, model.decoder.mapper(emb).view(emb.size(0), model.decoder.prefix_len, hid)


def pin_eos(out_dir):
    """The eos / finished branch of the reference loop (benchmark_baseline.py:212-224) on weights doctored so that it fires:
    rows that finish at step 0, rows that never finish, and a batch in which every row finishes at once (early break :224)."""
    from oracle import eos_fixture as EF
    a = synthetic.ARCHS["tiny"]
    model, sd = build_reference_model("tiny", 1234)
    EF.doctor(sd, a.gpt_dim)
    own = model.state_dict()
    load = dict(sd)
    for k in own:
        if k.startswith("encoder.backbone.model.heads."):
            load[k] = own[k]
    model.load_state_dict(load, strict=True)
    n_new, n_rows = 10, 32
    prefix = EF.prefixes(n_rows, a.prefix_len, a.gpt_dim)
    gpt2 = model.decoder.model

    @torch.inference_mode()
    def run(pre):
        B = pre.shape[0]
        prompt_ids = torch.tensor([[EF.PROMPT]], dtype=torch.long).expand(B, -1)
        full = torch.cat([pre, gpt2.transformer.wte(prompt_ids)], dim=1)
        mask = torch.ones(full.shape[:2], dtype=torch.long)
        toks = [[] for _ in range(B)]
        past, nxt_in = None, full
        finished = torch.zeros(B, dtype=torch.bool)
        margins, top2 = [], []
        steps_run = 0
        for _ in range(n_new):                      # benchmark_baseline.py:193-231 minus the CUDA events
            out = gpt2(inputs_embeds=nxt_in, attention_mask=mask, past_key_values=past, use_cache=True, return_dict=True)
            logits = out.logits[:, -1, :].float()
            steps_run += 1
            margins.append(EF.eos_margin(logits))
            t2 = logits.topk(2, dim=-1).values
            top2.append(t2[:, 0] - t2[:, 1])
            nxt = torch.argmax(logits, dim=-1)
            nxt = torch.where(finished, torch.full_like(nxt, EF.EOS), nxt)
            for i, t in enumerate(nxt.tolist()):
                if not finished[i]:
                    toks[i].append(t)
                    if t == EF.EOS:
                        finished[i] = True
            past = out.past_key_values
            if finished.all():
                break
            nxt_in = gpt2.transformer.wte(nxt).unsqueeze(1)
            mask = torch.cat([mask, torch.ones((B, 1), dtype=torch.long)], dim=1)
        ids = torch.full((B, n_new), EF.EOS, dtype=torch.int64)
        for i, t in enumerate(toks):
            ids[i, :len(t)] = torch.tensor(t)
        pad = n_new - len(margins)
        M = torch.stack(margins + [torch.zeros(B)] * pad, 0)
        T = torch.stack(top2 + [torch.zeros(B)] * pad, 0)
        return ids, torch.tensor([len(t) for t in toks]), M, T, steps_run

    ids, lens, M, T, steps = run(prefix)
    first = [i for i in range(n_rows) if lens[i] == 1]
    ids1, lens1, M1, T1, steps1 = run(prefix[first])
    print("eos fixture: lengths", lens.tolist(), "| all-finish-at-once batch rows", first, "steps run", steps1, flush=True)
    assert len(first) >= 3 and steps1 == 1 and (lens == n_new).any()
    np.savez_compressed(out_dir / "eos_tiny.npz", seed=1234, n_rows=n_rows, max_new_tokens=n_new, ids=ids.numpy(), lengths=lens.numpy(),
                        eos_margin=M.numpy(), top2_margin=T.numpy(), steps_run=steps, first_rows=np.array(first), first_ids=ids1.numpy(),
                        first_lengths=lens1.numpy(), first_steps_run=steps1)


def main():
    out_dir = REPO / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(8)
    if "--only-eos" in sys.argv:
        pin_eos(out_dir)
        return

    # ---- preprocessing: the reference's own load_video_tensor on JPEG files ----
    from PIL import Image
    from core.preprocessing.frame_loader import load_video_tensor
    with tempfile.TemporaryDirectory() as td:
        fr = synthetic.make_frames_u8(3, num_frames=10, size=224).numpy()
        for i in range(10):
            Image.fromarray(fr[i]).save(Path(td) / f"frame_{i:06d}.jpg", quality=95)
        ref = load_video_tensor(td, num_frames=4, image_size=224, device="cpu")[0]   # [4,3,224,224]
        picked = [0, 2, 4, 6]
        dec = np.stack([np.asarray(Image.open(Path(td) / f"frame_{i:06d}.jpg").convert("RGB")) for i in picked])
    # keep a crop of the decoded bytes + the reference's fp32 output for them
    np.savez_compressed(out_dir / "preprocess.npz", u8=dec[:, :32, :48, :], ref=ref[:, :, :32, :48].numpy(),
                        n_files=10, num_frames=4, picked=np.array(picked))

    # ---- tiny (2+2 layers) and full (12+12) models ----
    for arch_name, B, T, new_tok in (("tiny", 2, 2, 12), ("vit_b16_gpt2", 2, 4, 20)):
        seed = 1234
        model, sd = build_reference_model(arch_name, seed)
        frames = synthetic.make_batch_u8(0, B, T)
        video = torch.stack([
            torch.stack([torch.from_numpy(np.asarray(Image.fromarray(f.numpy()))).permute(2, 0, 1) for f in v])
            for v in frames]).float().div(255)
        mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225]).view(1, 1, 3, 1, 1)
        video = video.sub(mean).div(std)   # functional.to_tensor + functional.normalize
        with torch.inference_mode():
            feat = model.encoder(video)
            emb, prefix = align(model, feat)
            ids, lengths, logits = reference_greedy(model, prefix, new_tok)
            # the reference's own bf16 path (fp16 flag is the shipped one; bf16 is what the B200 build uses)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                feat_bf = model.encoder(video).float()
            _, prefix_bf = align(model, feat_bf)
            import copy
            model_bf = copy.deepcopy(model.decoder.model).to(torch.bfloat16)  # never round the fp32 model in place
            # teacher-forced bf16 logits with the fp32 ids
            gpt2 = model_bf
            full = torch.cat([prefix_bf.to(torch.bfloat16),
                              gpt2.transformer.wte(torch.tensor([[50256]]).expand(B, -1))], dim=1)
            past, x = None, full
            bf_logits = []
            for s in range(len(logits)):
                o = gpt2(inputs_embeds=x, past_key_values=past, use_cache=True, return_dict=True)
                bf_logits.append(o.logits[:, -1, :].float())
                past = o.past_key_values
                x = gpt2.transformer.wte(ids[:, s]).unsqueeze(1)
            del model_bf, gpt2
            beams = {}
            for nb, mx in ((3, 24), (5, 30)):
                if arch_name != "tiny" and nb == 3:
                    continue
                model.decoder.generate(emb, prompt="", max_new_tokens=mx, num_beams=nb, temperature=1.0, top_p=1.0,
                                       no_repeat_ngram_size=3, repetition_penalty=1.1)
                beams[f"beam{nb}_{mx}"] = _StubTok.captured.numpy()
            # HF greedy with processors (num_beams=1, temperature=1.0 => do_sample False)
            model.decoder.generate(emb, prompt="", max_new_tokens=new_tok, num_beams=1, temperature=1.0, top_p=1.0,
                                   no_repeat_ngram_size=3, repetition_penalty=1.1)
            hf_greedy = _StubTok.captured.numpy()
        L = torch.stack(logits, 0)            # [steps,B,V]
        Lb = torch.stack(bf_logits, 0)
        agree = (L.argmax(-1) == Lb.argmax(-1)).float().mean().item()
        top2 = L.topk(2, dim=-1).values
        stats = dict(
            feat_bf16_maxabs=(feat_bf - feat).abs().max().item(),
            feat_bf16_cos=torch.nn.functional.cosine_similarity(feat_bf, feat, dim=-1).min().item(),
            logits_bf16_maxabs=(Lb - L).abs().max().item(),
            logits_bf16_cos=torch.nn.functional.cosine_similarity(Lb, L, dim=-1).min().item(),
            logits_std=L.std().item(), tf_agree_bf16=agree,
            margin_median=(top2[..., 0] - top2[..., 1]).median().item(),
            margin_min=(top2[..., 0] - top2[..., 1]).min().item(),
        )
        print(arch_name, stats, "ids", ids.tolist(), flush=True)
        np.savez_compressed(
            out_dir / f"path_{arch_name}.npz", seed=seed, B=B, T=T, max_new_tokens=new_tok,
            feat=feat.numpy(), prefix=prefix.numpy(), ids=ids.numpy(), lengths=lengths.numpy(),
            logits_sub=L[:, :, ::LOGIT_STRIDE].numpy(), logits_stride=LOGIT_STRIDE,
            logits_top=L.topk(8, dim=-1).values.numpy(), logits_top_idx=L.topk(8, dim=-1).indices.numpy(),
            hf_greedy=hf_greedy, **beams, **{f"stat_{k}": v for k, v in stats.items()})
        del model
    pin_eos(out_dir)


if __name__ == "__main__":
    main()
