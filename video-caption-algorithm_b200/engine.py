"""Engine-level mirror of the reference's operator surface for this path.

reference                                        here
-----------------------------------------------  ---------------------------------------------
core/config.py:47-72      InferenceConfig        InferenceConfig (same field names; backend="b200")
core/inference.py:4-16    preset_to_kwargs       preset_to_kwargs (same presets)
core/models/model_loader.py:13-28  load_caption_model   load_caption_model (backend switch: "b200")
core/engine.py:66-83      InferenceEngine.infer  InferenceEngine.infer_frames / infer

Differences that are deliberate: the ViT encode runs ONCE per request and its
prefix is reused by the three candidates (the reference re-runs the encoder inside
`_generate_once`, engine.py:43); frames arrive as uint8 tensors (or a directory of already
image_size x image_size frames), the JPEG decode + PIL resize front end stays with the caller.
Host string work (`clean_text`, `select_best`, core/postprocessing/*) is out of scope and can be
passed in as callables.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path
from typing import Callable, Optional

import torch

from .model import B200CaptionModel, EOS


@dataclass(frozen=True)
class InferenceConfig:
    """Fields of core/config.py:47-72 that reach this path (the torch.compile / channels_last / CuPy /
    TensorRT switches of the reference have no meaning here and are accepted but ignored)."""
    ckpt: str = ""
    vit_name: str = "vit_base_patch16_224"
    gpt2_name: str = "gpt2"
    prefix_len: int = 4
    num_frames: int = 8
    image_size: int = 224
    ln_scale: float = 0.6
    in_weight: float = 0.4
    preset1: str = "precise"
    preset2: str = "precise"
    preset3: str = "natural"
    prompt1: str = ""
    prompt2: str = "State the main action in one short sentence:"
    prompt3: str = "Write a short, natural caption:"
    device: str = "cuda:0"
    backend: str = "b200"
    enable_mlp_bias_gelu_fusion: bool = True      # ViTOptimizeConfig: selects tanh GELU on the timm layout (video_encoder.py:123-134)
    ignored: dict = field(default_factory=dict)


_HEADS = {"vit_base_patch16_224": 12, "vit_large_patch14_224": 16, "gpt2": 12, "gpt2-medium": 16}


def preset_to_kwargs(name: str) -> dict:
    """Decode-policy registry, same numbers as core/inference.py:4-16."""
    name = (name or "precise").lower()
    table = {
        "precise": dict(num_beams=3, max_new_tokens=24, temperature=1.0, top_p=1.0, no_repeat_ngram_size=3, repetition_penalty=1.1),
        "detailed": dict(num_beams=4, max_new_tokens=40, temperature=1.0, top_p=1.0, no_repeat_ngram_size=3, repetition_penalty=1.1),
        "natural": dict(num_beams=1, max_new_tokens=24, temperature=0.9, top_p=0.9, no_repeat_ngram_size=3, repetition_penalty=1.05),
        "safe_sample": dict(num_beams=1, max_new_tokens=22, temperature=0.8, top_p=0.85, no_repeat_ngram_size=3, repetition_penalty=1.1),
    }
    return dict(table.get(name, table["precise"]))


def read_checkpoint(path) -> dict:
    """core/models/model_loader.py:31-40, :73-75: a checkpoint file is either a raw state-dict or {"model_state": state-dict,
    ...}; tensors only (loaded with weights_only=True — no pickled code is executed, unlike the reference's loader)."""
    state = torch.load(Path(path), map_location="cpu", weights_only=True)
    if isinstance(state, dict) and "model_state" in state:
        state = state["model_state"]
    if not isinstance(state, dict) or not all(isinstance(v, torch.Tensor) for v in state.values()):
        raise ValueError(f"{path}: expected a state-dict of tensors (optionally under 'model_state')")
    return state


def load_caption_model(config: InferenceConfig, state_dict: Optional[dict] = None, tokenizer=None) -> B200CaptionModel:
    """The backend switch of core/models/model_loader.py:21-28 with one more value, "b200".
    `state_dict` may be passed directly (tests, benchmark); otherwise `config.ckpt` is loaded like
    model_loader.py:31-40 / :73-75 ({"model_state": …} or a raw state-dict)."""
    backend = config.backend.lower()
    if backend != "b200":
        raise ValueError(f"this package only provides backend='b200' (got {config.backend!r}); the torch backend is the reference itself")
    if state_dict is None:
        if not config.ckpt:
            raise FileNotFoundError("InferenceConfig.ckpt is empty and no state_dict was given")
        state_dict = read_checkpoint(config.ckpt)
    gelu = None if config.enable_mlp_bias_gelu_fusion else "erf"
    return B200CaptionModel(state_dict, config.device, vit_heads=_HEADS.get(config.vit_name, 12), gpt_heads=_HEADS.get(config.gpt2_name, 12),
                            gelu=gelu, tokenizer=tokenizer, ln_scale=config.ln_scale, in_weight=config.in_weight)


def sample_frame_indices(n_files: int, num_frames: int) -> list[int]:
    """core/preprocessing/frame_loader.py:31-32: step = max(n // T, 1); files[::step][:T] (no padding)."""
    step = max(n_files // num_frames, 1)
    return list(range(0, n_files, step))[:num_frames]


class InferenceEngine:
    """Stateless engine (core/engine.py:20-37): owns the model and the tensor flow only."""

    def __init__(self, config: InferenceConfig, state_dict: Optional[dict] = None, tokenizer=None,
                 clean_text: Optional[Callable[[str], str]] = None, select_best: Optional[Callable] = None):
        self.config = config
        self.model = load_caption_model(config, state_dict, tokenizer)
        # core/engine.py:75-83: clean_text on every candidate, select_best over the three (host string work; postprocess.py)
        from . import postprocess
        self._clean = clean_text or postprocess.clean_text
        self._select = select_best or postprocess.select_best

    @classmethod
    def from_config(cls, config: InferenceConfig, **kw):
        return cls(config, **kw)

    def load_frames_dir(self, frames_dir: str) -> torch.Tensor:
        """frame_loader.py:13-47: sample `num_frames` of the frame_*.jpg files, decode (PIL, host), and return uint8
        [1,T,H,W,3] on the device at the files' own size; `encode_prefix` then applies the Pillow-exact resize to
        image_size x image_size on the GPU (resample.py), ToTensor/Normalize and the patch layout."""
        from .frames import FrameDecodePool
        if getattr(self, "_decode_pool", None) is None:
            self._decode_pool = FrameDecodePool()          # PIL decode on a pool of host threads into pinned memory (frames.py)
        return self._decode_pool.decode_dirs([frames_dir], self.config.num_frames).to(self.model.device, non_blocking=True)

    @torch.no_grad()
    def infer_frames(self, frames_u8: torch.Tensor) -> dict:
        """uint8 [B,T,H,W,3] -> per-candidate token ids (and texts when a tokenizer is attached)."""
        m, cfg = self.model, self.config
        feat, _ = m.encode_prefix(frames_u8)                       # once, not three times (engine.py:43)
        emb = self._prefix_embedding(feat)
        out = {}
        for key, prompt, preset in (("S1", cfg.prompt1, cfg.preset1), ("S2", cfg.prompt2, cfg.preset2), ("S3", cfg.prompt3, cfg.preset3)):
            kw = preset_to_kwargs(preset)
            # nucleus-sampling presets (temperature != 1) draw from torch's CUDA generator on the device: reproducible under
            # torch.manual_seed, excluded from token-for-token parity
            if prompt and m.decoder.tokenizer is None:
                prompt = ""                                        # no BPE files offline: fall back to the bos prompt
            texts = m.decoder.generate(emb, prompt=prompt, **kw)
            out[key] = dict(ids=m.decoder.last_ids, lengths=m.decoder.last_lengths, text=[self._clean(t) for t in texts])
        if self._select is not None and m.decoder.tokenizer is not None:
            out["BEST"] = [self._select([(k, out[k]["text"][b]) for k in ("S1", "S2", "S3")]) for b in range(frames_u8.shape[0])]
        return out

    def _prefix_embedding(self, feat: torch.Tensor) -> torch.Tensor:
        """engine.py:44-50 as the reference writes it (proj=Identity, unsqueeze, layer_norm*ln_scale, *in_weight):
        [B,256] -> [B,1,256].  Tiny host-side glue kept in torch; the fused kernel path is `encode_prefix`."""
        emb = feat.unsqueeze(1)
        if self.config.ln_scale and self.config.ln_scale > 0:
            emb = torch.nn.functional.layer_norm(emb, emb.shape[-1:]) * self.config.ln_scale
        if self.config.in_weight and self.config.in_weight > 0:
            emb = emb * self.config.in_weight
        return emb

    def infer(self, frames_dir: str) -> dict:
        return self.infer_frames(self.load_frames_dir(frames_dir))
