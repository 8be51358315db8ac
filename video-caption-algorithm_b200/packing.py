"""state_dict -> device weight blobs in the layout the kernels want.

Accepts the two key layouts the reference can produce:
torchvision fallback (`encoder.backbone.model.*`, src/models/video_encoder.py:84-103)
and timm (`encoder.backbone.blocks.N.*`, video_encoder.py:69-80), plus HF GPT-2
(`decoder.model.transformer.*`, Conv1D weights stored [in,out]) and the
`decoder.mapper.0` / `encoder.proj` linears, as saved by
core/models/model_loader.py:73-81 (`{"model_state": …}` or a raw state-dict).

GEMM operands become bf16 exactly as `tensor.to(torch.bfloat16)` (round to nearest
even); biases, LayerNorm affines, position tables and the two tiny linears of the
alignment stage stay fp32.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import lib as L


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


@dataclass
class PackedModel:
    vit: L.VcVitWeights
    gpt: L.VcGptWeights
    mapper_w: torch.Tensor
    mapper_b: torch.Tensor
    lut: torch.Tensor
    keep: list          # owns every device tensor and ctypes array referenced by the structs
    dims: dict


def unwrap_state(state: dict) -> dict:
    """model_loader.py:74-75: checkpoints are either {"model_state": sd} or sd."""
    if isinstance(state, dict) and "model_state" in state:
        return state["model_state"]
    return state


def fold_layernorm(w: torch.Tensor, bias: torch.Tensor | None, gamma: torch.Tensor, beta: torch.Tensor):
    """LayerNorm folded into the product that consumes it:  LN(x) W^T + b = rstd * (x W'^T - mean * cs) + b'
    with W' = bf16(gamma (.) W) [out, in] (what the kernel multiplies), cs[n] = sum_k W'[n][k] (of the ROUNDED weights, so the
    mean term cancels exactly what the product accumulated) and b' = b + W beta (fp32, from the unrounded weights)."""
    w = w.detach().double()
    wf = (w * gamma.detach().double().view(1, -1)).float().to(torch.bfloat16)
    cs = wf.double().sum(dim=1).float()
    b2 = w @ beta.detach().double()
    if bias is not None:
        b2 = b2 + bias.detach().double()
    return wf, cs, b2.float()


def pack(state: dict, device: torch.device, *, vit_heads: int, gpt_heads: int, gelu: str | None = None) -> PackedModel:
    sd = unwrap_state(state)
    keep: list = []
    raw = lambda t: _own(keep, t.to(device=device).contiguous())
    bf = lambda t: _own(keep, t.detach().to(device=device, dtype=torch.float32).to(torch.bfloat16).contiguous())
    f32 = lambda t: _own(keep, t.detach().to(device=device, dtype=torch.float32).contiguous())

    # ------------------------------------------------------------- ViT
    if "encoder.backbone.model.class_token" in sd:
        p = "encoder.backbone.model."
        names = dict(cls=p + "class_token", pos=p + "encoder.pos_embedding", pw=p + "conv_proj.weight", pb=p + "conv_proj.bias",
                     lnf=p + "encoder.ln", ln1="ln_1", qkv_w="self_attention.in_proj_weight", qkv_b="self_attention.in_proj_bias",
                     out="self_attention.out_proj", ln2="ln_2", fc1="mlp.0", fc2="mlp.3")
        blk = lambda i: f"{p}encoder.layers.encoder_layer_{i}."
        n_layers = sum(1 for k in sd if k.startswith(p + "encoder.layers.") and k.endswith("ln_1.weight"))
        default_gelu = "erf"       # torchvision nn.GELU()
    elif "encoder.backbone.cls_token" in sd:
        p = "encoder.backbone."
        names = dict(cls=p + "cls_token", pos=p + "pos_embed", pw=p + "patch_embed.proj.weight", pb=p + "patch_embed.proj.bias",
                     lnf=p + "norm", ln1="norm1", qkv_w="attn.qkv.weight", qkv_b="attn.qkv.bias", out="attn.proj", ln2="norm2",
                     fc1="mlp.fc1", fc2="mlp.fc2")
        blk = lambda i: f"{p}blocks.{i}."
        n_layers = sum(1 for k in sd if k.startswith(p + "blocks.") and k.endswith("norm1.weight"))
        default_gelu = "tanh"      # video_encoder.py:123-134 flips timm's GELU to the tanh form
    else:
        raise KeyError("state_dict holds neither the torchvision nor the timm ViT key layout")
    gelu = gelu or default_gelu
    pw = sd[names["pw"]]
    D, patch = pw.shape[0], pw.shape[-1]
    k_raw = 3 * patch * patch
    k_pad = _round_up(k_raw, 64)
    pw2 = torch.zeros(D, k_pad, dtype=torch.float32)
    pw2[:, :k_raw] = pw.reshape(D, k_raw).float()          # columns ordered (c, i, j) like Conv2d
    pos = sd[names["pos"]].reshape(-1, D).float()
    tokens = pos.shape[0]
    cls_pos0 = sd[names["cls"]].reshape(D).float() + pos[0]
    layers = (L.VcVitLayer * n_layers)()
    mlp = sd[blk(0) + names["fc1"] + ".weight"].shape[0]
    for i in range(n_layers):
        b = blk(i)
        ly = layers[i]
        ly.ln1_g, ly.ln1_b = f32(sd[b + names["ln1"] + ".weight"]), f32(sd[b + names["ln1"] + ".bias"])
        ly.qkv_w, ly.qkv_b = bf(sd[b + names["qkv_w"]]), f32(sd[b + names["qkv_b"]])
        ly.proj_w, ly.proj_b = bf(sd[b + names["out"] + ".weight"]), f32(sd[b + names["out"] + ".bias"])
        ly.ln2_g, ly.ln2_b = f32(sd[b + names["ln2"] + ".weight"]), f32(sd[b + names["ln2"] + ".bias"])
        ly.fc1_w, ly.fc1_b = bf(sd[b + names["fc1"] + ".weight"]), f32(sd[b + names["fc1"] + ".bias"])
        ly.fc2_w, ly.fc2_b = bf(sd[b + names["fc2"] + ".weight"]), f32(sd[b + names["fc2"] + ".bias"])
        # LayerNorm folded into the consumer GEMM (csrc/gemm_tcgen05.cu VC_EPI_LNF_*): no stand-alone LayerNorm pass in the encoder
        wf, cs, b2 = fold_layernorm(sd[b + names["qkv_w"]], sd[b + names["qkv_b"]], sd[b + names["ln1"] + ".weight"], sd[b + names["ln1"] + ".bias"])
        ly.qkv_wf, ly.qkv_cs, ly.qkv_bf = raw(wf), raw(cs), raw(b2)
        wf, cs, b2 = fold_layernorm(sd[b + names["fc1"] + ".weight"], sd[b + names["fc1"] + ".bias"], sd[b + names["ln2"] + ".weight"],
                                    sd[b + names["ln2"] + ".bias"])
        ly.fc1_wf, ly.fc1_cs, ly.fc1_bf = raw(wf), raw(cs), raw(b2)
    keep.append(layers)
    video_dim = sd["encoder.proj.weight"].shape[0]
    vit = L.VcVitWeights()
    vit.dim, vit.layers, vit.heads, vit.mlp, vit.tokens, vit.patch_k = D, n_layers, vit_heads, mlp, tokens, k_pad
    vit.gelu_tanh = 1 if gelu == "tanh" else 0
    vit.video_dim = video_dim
    vit.patch_w, vit.patch_b = bf(pw2), f32(sd[names["pb"]])
    vit.cls_pos0, vit.pos = f32(cls_pos0), f32(pos)
    vit.lnf_g, vit.lnf_b = f32(sd[names["lnf"] + ".weight"]), f32(sd[names["lnf"] + ".bias"])
    vit.head_w, vit.head_b = f32(sd["encoder.proj.weight"]), f32(sd["encoder.proj.bias"])
    vit.layer = C.cast(layers, C.POINTER(L.VcVitLayer))

    # ------------------------------------------------------------- GPT-2
    g = "decoder.model.transformer."
    wte = sd[g + "wte.weight"].float()
    vocab, H = wte.shape
    vocab_pad = _round_up(vocab, 256)
    wte_p = torch.zeros(vocab_pad, H, dtype=torch.float32)
    wte_p[:vocab] = wte
    n_gl = sum(1 for k in sd if k.startswith(g + "h.") and k.endswith("ln_1.weight"))
    glayers = (L.VcGptLayer * n_gl)()
    for i in range(n_gl):
        b = f"{g}h.{i}."
        ly = glayers[i]
        ly.ln1_g, ly.ln1_b = f32(sd[b + "ln_1.weight"]), f32(sd[b + "ln_1.bias"])
        # HF Conv1D weight is [in,out]; the GEMM's B operand is [out,in] K-major
        ly.attn_w, ly.attn_b = bf(sd[b + "attn.c_attn.weight"].t()), f32(sd[b + "attn.c_attn.bias"])
        ly.aproj_w, ly.aproj_b = bf(sd[b + "attn.c_proj.weight"].t()), f32(sd[b + "attn.c_proj.bias"])
        ly.ln2_g, ly.ln2_b = f32(sd[b + "ln_2.weight"]), f32(sd[b + "ln_2.bias"])
        ly.fc_w, ly.fc_b = bf(sd[b + "mlp.c_fc.weight"].t()), f32(sd[b + "mlp.c_fc.bias"])
        ly.mproj_w, ly.mproj_b = bf(sd[b + "mlp.c_proj.weight"].t()), f32(sd[b + "mlp.c_proj.bias"])
        # decode chain: ln_1 folded into c_attn, ln_2 into c_fc (csrc/decode_chain.cu)
        wf, cs, b2 = fold_layernorm(sd[b + "attn.c_attn.weight"].t(), sd[b + "attn.c_attn.bias"], sd[b + "ln_1.weight"], sd[b + "ln_1.bias"])
        ly.attn_wf, ly.attn_cs, ly.attn_bf = raw(wf), raw(cs), raw(b2)
        wf, cs, b2 = fold_layernorm(sd[b + "mlp.c_fc.weight"].t(), sd[b + "mlp.c_fc.bias"], sd[b + "ln_2.weight"], sd[b + "ln_2.bias"])
        ly.fc_wf, ly.fc_cs, ly.fc_bf = raw(wf), raw(cs), raw(b2)
    keep.append(glayers)
    gpt = L.VcGptWeights()
    gpt.dim, gpt.layers, gpt.heads, gpt.vocab, gpt.vocab_pad = H, n_gl, gpt_heads, vocab, vocab_pad
    gpt.n_pos = sd[g + "wpe.weight"].shape[0]
    gpt.wte, gpt.wpe = bf(wte_p), f32(sd[g + "wpe.weight"])
    gpt.lnf_g, gpt.lnf_b = f32(sd[g + "ln_f.weight"]), f32(sd[g + "ln_f.bias"])
    wf, cs, b2 = fold_layernorm(wte_p, None, sd[g + "ln_f.weight"], sd[g + "ln_f.bias"])      # ln_f folded into the tied lm_head
    gpt.lmh_w, gpt.lmh_cs, gpt.lmh_b = raw(wf), raw(cs), raw(b2)
    gpt.layer = C.cast(glayers, C.POINTER(L.VcGptLayer))

    mapper_w = sd["decoder.mapper.0.weight"].detach().to(device=device, dtype=torch.float32).contiguous()
    mapper_b = sd["decoder.mapper.0.bias"].detach().to(device=device, dtype=torch.float32).contiguous()
    lut = normalize_lut().to(device)
    dims = dict(vit_dim=D, patch=patch, k_pad=k_pad, tokens=tokens, vit_layers=n_layers, vit_heads=vit_heads, mlp=mlp,
                video_dim=video_dim, gpt_dim=H, gpt_layers=n_gl, gpt_heads=gpt_heads, vocab=vocab, vocab_pad=vocab_pad,
                prefix_len=mapper_w.shape[0] // H, n_pos=int(gpt.n_pos), gelu=gelu)
    return PackedModel(vit, gpt, mapper_w, mapper_b, lut, keep, dims)


def _own(keep: list, t: torch.Tensor) -> int:
    keep.append(t)
    return t.data_ptr()


def normalize_lut() -> torch.Tensor:
    """[3,256] fp32: ToTensor (`u8.float().div(255)`) then Normalize (`.sub(mean).div(std)`),
    built with the reference's op sequence (core/preprocessing/frame_loader.py:34-40) so the
    kernel's table lookup reproduces it bit for bit."""
    v = torch.arange(256, dtype=torch.uint8).to(torch.float32).div(255)
    mean = torch.tensor([0.485, 0.456, 0.406], dtype=torch.float32).view(3, 1)
    std = torch.tensor([0.229, 0.224, 0.225], dtype=torch.float32).view(3, 1)
    return v.view(1, 256).sub(mean).div(std).contiguous()
