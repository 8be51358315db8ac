"""Deterministic synthetic weights and frames.

The reference ships no checkpoint and needs the network to build its model
(`pretrained=True`, src/models/caption_model.py:44; `from_pretrained`,
src/models/text_decoder.py:27-28), so every parity test and the benchmark use a
random-init state-dict in the reference's own key layout.
Each tensor is drawn from a generator seeded by (seed, crc32(key)), so any
subset (a 2-layer test model, one shard of a big one) reproduces bit-for-bit
on any machine with the same torch build, without shipping a file.

Pure host code: torch CPU only, no CUDA, no oracle import.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass

import torch


@dataclass(frozen=True)
class Arch:
    """Shapes of one ViT + GPT-2 pair."""

    name: str
    image: int
    patch: int
    vit_dim: int
    vit_layers: int
    vit_heads: int
    vit_mlp: int
    video_dim: int
    gpt_dim: int
    gpt_layers: int
    gpt_heads: int
    vocab: int
    n_pos: int
    prefix_len: int

    @property
    def grid(self) -> int:
        return self.image // self.patch

    @property
    def tokens(self) -> int:
        return self.grid * self.grid + 1

    @property
    def patch_k(self) -> int:
        return 3 * self.patch * self.patch


ARCHS = {
    # ViT-B/16 + GPT-2 small: BASELINE.json configs[0..3]
    "vit_b16_gpt2": Arch("vit_b16_gpt2", 224, 16, 768, 12, 12, 3072, 256, 768, 12, 12, 50257, 1024, 4),
    # ViT-L/14 + GPT-2 medium: BASELINE.json configs[4]
    "vit_l14_gpt2m": Arch("vit_l14_gpt2m", 224, 14, 1024, 24, 16, 4096, 256, 1024, 24, 16, 50257, 1024, 4),
    # 2-layer cut of the B/16 pair: same widths, fast enough for CPU-side tests
    "tiny": Arch("tiny", 224, 16, 768, 2, 12, 3072, 256, 768, 2, 12, 50257, 1024, 4),
    # 2-layer cut of the L/14 + medium pair (patch K=588 padded to 640, 257 tokens, 16 heads, width 1024)
    "tiny_l14": Arch("tiny_l14", 224, 14, 1024, 2, 16, 4096, 256, 1024, 2, 16, 50257, 1024, 4),
}


def _draw(seed: int, key: str, shape, std: float, mean: float = 0.0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1_000_003 + zlib.crc32(key.encode())) % (2**63 - 1))
    t = torch.randn(tuple(shape), generator=g, dtype=torch.float32)
    return t.mul_(std).add_(mean)


def make_state_dict(arch: Arch | str = "vit_b16_gpt2", seed: int = 1234, layout: str = "torchvision") -> dict:
    """fp32 state-dict in the reference model's key layout.

    layout="torchvision": keys of the reference's torchvision fallback wrapper
    (src/models/video_encoder.py:84-103) — `encoder.backbone.model.*`.
    layout="timm": keys of the timm backbone (video_encoder.py:69-80) —
    `encoder.backbone.blocks.N.*`.  Same values, different names.
    Biases and LayerNorm affines are non-trivial on purpose so that every fused
    epilogue is exercised (the stock inits are zeros/ones).
    """
    a = ARCHS[arch] if isinstance(arch, str) else arch
    D, Dm, L = a.vit_dim, a.vit_mlp, a.vit_layers
    sd: dict[str, torch.Tensor] = {}

    def put(key_tv: str, key_timm: str, shape, std, mean=0.0):
        key = key_tv if layout == "torchvision" else key_timm
        # always seed from the torchvision name so both layouts hold equal values
        sd[key] = _draw(seed, key_tv, shape, std, mean)

    tv = "encoder.backbone.model."
    tm = "encoder.backbone."
    put(tv + "class_token", tm + "cls_token", (1, 1, D), 0.02)
    put(tv + "conv_proj.weight", tm + "patch_embed.proj.weight", (D, 3, a.patch, a.patch), (1.0 / a.patch_k) ** 0.5)
    put(tv + "conv_proj.bias", tm + "patch_embed.proj.bias", (D,), 0.02)
    put(tv + "encoder.pos_embedding", tm + "pos_embed", (1, a.tokens, D), 0.02)
    for i in range(L):
        p = f"{tv}encoder.layers.encoder_layer_{i}."
        q = f"{tm}blocks.{i}."
        put(p + "ln_1.weight", q + "norm1.weight", (D,), 0.1, 1.0)
        put(p + "ln_1.bias", q + "norm1.bias", (D,), 0.05)
        put(p + "self_attention.in_proj_weight", q + "attn.qkv.weight", (3 * D, D), (2.0 / (4 * D)) ** 0.5)
        put(p + "self_attention.in_proj_bias", q + "attn.qkv.bias", (3 * D,), 0.02)
        put(p + "self_attention.out_proj.weight", q + "attn.proj.weight", (D, D), (1.0 / (3 * D)) ** 0.5)
        put(p + "self_attention.out_proj.bias", q + "attn.proj.bias", (D,), 0.02)
        put(p + "ln_2.weight", q + "norm2.weight", (D,), 0.1, 1.0)
        put(p + "ln_2.bias", q + "norm2.bias", (D,), 0.05)
        put(p + "mlp.0.weight", q + "mlp.fc1.weight", (Dm, D), (2.0 / (D + Dm)) ** 0.5)
        put(p + "mlp.0.bias", q + "mlp.fc1.bias", (Dm,), 0.02)
        put(p + "mlp.3.weight", q + "mlp.fc2.weight", (D, Dm), (2.0 / (D + Dm)) ** 0.5)
        put(p + "mlp.3.bias", q + "mlp.fc2.bias", (D,), 0.02)
    put(tv + "encoder.ln.weight", tm + "norm.weight", (D,), 0.1, 1.0)
    put(tv + "encoder.ln.bias", tm + "norm.bias", (D,), 0.05)

    sd["encoder.proj.weight"] = _draw(seed, "encoder.proj.weight", (a.video_dim, D), (1.0 / (3 * D)) ** 0.5)
    sd["encoder.proj.bias"] = _draw(seed, "encoder.proj.bias", (a.video_dim,), 0.02)

    H, Lg = a.gpt_dim, a.gpt_layers
    g = "decoder.model.transformer."
    sd[g + "wte.weight"] = _draw(seed, g + "wte.weight", (a.vocab, H), 0.02)
    sd[g + "wpe.weight"] = _draw(seed, g + "wpe.weight", (a.n_pos, H), 0.02)
    for i in range(Lg):
        p = f"{g}h.{i}."
        sd[p + "ln_1.weight"] = _draw(seed, p + "ln_1.weight", (H,), 0.1, 1.0)
        sd[p + "ln_1.bias"] = _draw(seed, p + "ln_1.bias", (H,), 0.05)
        # HF Conv1D stores [in, out] (modeling_gpt2 Conv1D: y = x @ W + b)
        sd[p + "attn.c_attn.weight"] = _draw(seed, p + "attn.c_attn.weight", (H, 3 * H), 0.02)
        sd[p + "attn.c_attn.bias"] = _draw(seed, p + "attn.c_attn.bias", (3 * H,), 0.02)
        sd[p + "attn.c_proj.weight"] = _draw(seed, p + "attn.c_proj.weight", (H, H), 0.02 / (2 * Lg) ** 0.5)
        sd[p + "attn.c_proj.bias"] = _draw(seed, p + "attn.c_proj.bias", (H,), 0.02)
        sd[p + "ln_2.weight"] = _draw(seed, p + "ln_2.weight", (H,), 0.1, 1.0)
        sd[p + "ln_2.bias"] = _draw(seed, p + "ln_2.bias", (H,), 0.05)
        sd[p + "mlp.c_fc.weight"] = _draw(seed, p + "mlp.c_fc.weight", (H, 4 * H), 0.02)
        sd[p + "mlp.c_fc.bias"] = _draw(seed, p + "mlp.c_fc.bias", (4 * H,), 0.02)
        sd[p + "mlp.c_proj.weight"] = _draw(seed, p + "mlp.c_proj.weight", (4 * H, H), 0.02 / (2 * Lg) ** 0.5)
        sd[p + "mlp.c_proj.bias"] = _draw(seed, p + "mlp.c_proj.bias", (H,), 0.02)
    sd[g + "ln_f.weight"] = _draw(seed, g + "ln_f.weight", (H,), 0.1, 1.0)
    sd[g + "ln_f.bias"] = _draw(seed, g + "ln_f.bias", (H,), 0.05)
    sd["decoder.model.lm_head.weight"] = sd[g + "wte.weight"]  # tied (HF GPT2LMHeadModel)
    sd["decoder.mapper.0.weight"] = _draw(seed, "decoder.mapper.0.weight", (H * a.prefix_len, a.video_dim), (1.0 / (3 * a.video_dim)) ** 0.5)
    sd["decoder.mapper.0.bias"] = _draw(seed, "decoder.mapper.0.bias", (H * a.prefix_len,), 0.02)
    return sd


def make_frames_u8(video_index: int, num_frames: int = 16, size: int = 224) -> torch.Tensor:
    """uint8 [T,H,W,3] frames of one video, reproducible per GLOBAL video index.

    Structured on purpose: per-video base colour, per-frame
    vertical ramp, and noise — plain noise makes all videos' features ~equal,
    which would hide a mis-sharded batch.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(10_000 + int(video_index))
    base = torch.randint(0, 256, (1, 1, 1, 3), generator=g).float() * 0.5
    phase = torch.randint(0, size, (num_frames, 1, 1, 1), generator=g).float()
    rows = torch.arange(size, dtype=torch.float32).view(1, size, 1, 1)
    amp = torch.randint(32, 129, (num_frames, 1, 1, 1), generator=g).float()
    ramp = ((rows + phase) % size) / size * amp
    noise = torch.randint(0, 64, (num_frames, size, size, 3), generator=g).float()
    return (base + ramp + noise).clamp_(0, 255).to(torch.uint8)


def make_batch_u8(first_video: int, n_videos: int, num_frames: int = 16, size: int = 224) -> torch.Tensor:
    """uint8 [B,T,H,W,3] for global video indices first_video .. first_video+n-1."""
    return torch.stack([make_frames_u8(first_video + i, num_frames, size) for i in range(n_videos)], 0)
