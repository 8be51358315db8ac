"""`B200CaptionModel`: the reference's model object, re-hosted on libvcb200.

It presents the module-attribute surface every reference caller uses
(call sites core/engine.py:43-61,
core/scripts/benchmark_baseline.py:162-289, core/scripts/profile_nsight.py:64-126):

    model.encoder(video)            -> [B, 256] fp32          (src/models/video_encoder.py:288)
    model.proj(x)                   -> x                      (caption_model.py:67, Identity)
    model.decoder.mapper(emb)       -> [..., P*H] fp32        (text_decoder.py:36-45)
    model.decoder.prefix_len / .cond_mode / .tokenizer
    model.decoder.model(inputs_embeds=…, past_key_values=…, use_cache=True) -> .logits, .past_key_values
    model.decoder.model.transformer.wte(ids), model.decoder.model.config.n_embd
    model.decoder.generate(emb, prompt=…, max_new_tokens=…, num_beams=…, …) -> List[str]

plus the fused fast path the B200 build adds: `caption_ids(frames_u8, …)` —
uint8 frames in, token ids out, no host sync inside.

torch is used here for device memory, streams and CUDA-graph capture only; all
arithmetic runs in csrc/ through the C ABI.  No CPU path: every entry point
raises if CUDA or the library is missing.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import os

import torch

from . import lib as L
from .resample import FrameResizer
from .memory import KvCache, Workspace
from .packing import PackedModel, pack

EOS = 50256  # GPT-2 bos = eos = pad id (text_decoder.py:29-30,122)


def _require_cuda(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise L.VcError("the b200 backend needs a CUDA device (there is no CPU fallback)")
    return dev


class _Encoder:
    """Callable standing where `ViTFrameEncoder` stands (video_encoder.py:288-326)."""

    def __init__(self, owner: "B200CaptionModel"):
        self._m = owner

    def __call__(self, video: torch.Tensor) -> torch.Tensor:
        with torch.cuda.device(self._m.device):       # kernels, func attributes and streams belong to the MODEL's device
            return self._call(video)

    def _call(self, video: torch.Tensor) -> torch.Tensor:
        m = self._m
        if video.dtype == torch.uint8:
            feat, _ = m.encode_prefix(video)
            return feat
        # the reference contract: normalised fp32 [B,T,3,H,W] (or [B,3,H,W])
        if video.ndim == 4:
            video = video.unsqueeze(1)
        if video.ndim != 5:
            raise ValueError(f"expect [B,T,3,H,W], got {tuple(video.shape)}")
        B, T, Cc, H, W = video.shape
        d = m.dims
        video = video.to(device=m.device, dtype=torch.float32).contiguous()
        n = B * T
        if n == 0:
            return torch.zeros(B, d["video_dim"], device=m.device, dtype=torch.float32)
        patches = m.ws.patches(n)
        lib = L.load()
        L.check(lib.vc_patchify_f32(video.data_ptr(), patches.data_ptr(), n, H, W, d["patch"], d["k_pad"], L.current_stream(m.device)))
        cls = m._encode_patches(patches, n)
        feat = torch.empty(B, d["video_dim"], device=m.device, dtype=torch.float32)
        L.check(lib.vc_vit_pool_temporal(cls.data_ptr(), 0, B, T, 1, d["vit_dim"], 0, m._pooled(B).data_ptr(), L.current_stream(m.device)))
        L.check(lib.vc_linear_bias_f32(m._pooled(B).data_ptr(), m.packed.vit.head_w, m.packed.vit.head_b, feat.data_ptr(), B,
                                       d["vit_dim"], d["video_dim"], L.current_stream(m.device)))
        return feat


class _Mapper:
    """`decoder.mapper` = Sequential(Linear(video_dim, H*P), Dropout) in eval mode (text_decoder.py:36-45).
    Same contract as `CuPyLinearCompat.forward` (cupy_linear_mapper.py:154-184) minus the fallback."""

    def __init__(self, owner: "B200CaptionModel"):
        self._m = owner
        self.last_backend = "b200"
        self.last_error = ""

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        with torch.cuda.device(self._m.device):
            return self._call(x)

    def _call(self, x: torch.Tensor) -> torch.Tensor:
        m = self._m
        shape = x.shape
        x2 = x.to(device=m.device, dtype=torch.float32).reshape(-1, shape[-1]).contiguous()
        out_f = m.packed.mapper_w.shape[0]
        y = torch.empty(x2.shape[0], out_f, device=m.device, dtype=torch.float32)
        L.check(L.load().vc_linear_bias_f32(x2.data_ptr(), m.packed.mapper_w.data_ptr(), m.packed.mapper_b.data_ptr(), y.data_ptr(),
                                            x2.shape[0], x2.shape[1], out_f, L.current_stream(m.device)))
        return y.reshape(*shape[:-1], out_f)


@dataclass
class _GptOutput:
    logits: torch.Tensor
    past_key_values: "KvCache"


class _Wte:
    def __init__(self, owner):
        self._m = owner

    def __call__(self, ids: torch.Tensor) -> torch.Tensor:
        with torch.cuda.device(self._m.device):
            return self._call(ids)

    def _call(self, ids: torch.Tensor) -> torch.Tensor:
        m = self._m
        flat = ids.to(device=m.device, dtype=torch.int32).reshape(-1).contiguous()
        out = torch.empty(flat.numel(), m.dims["gpt_dim"], device=m.device, dtype=torch.float32)
        L.check(L.load().vc_gpt2_embed_tokens(C.byref(m.packed.gpt), flat.data_ptr(), flat.numel(), out.data_ptr(), L.current_stream(m.device)))
        return out.reshape(*ids.shape, m.dims["gpt_dim"])


@dataclass
class _Cfg:
    n_embd: int
    n_layer: int
    n_head: int
    vocab_size: int
    n_positions: int


class _Transformer:
    def __init__(self, owner):
        self.wte = _Wte(owner)


class _Gpt2:
    """Stands where HF `GPT2LMHeadModel` stands in benchmark_baseline.py:162-231:
    `gpt2(inputs_embeds=…, attention_mask=…, past_key_values=…, use_cache=True, return_dict=True)`.
    `attention_mask` must be all ones (it always is on this path: text_decoder.py:125-129)."""

    def __init__(self, owner: "B200CaptionModel"):
        self._m = owner
        d = owner.dims
        self.config = _Cfg(d["gpt_dim"], d["gpt_layers"], d["gpt_heads"], d["vocab"], d["n_pos"])
        self.transformer = _Transformer(owner)

    def __call__(self, inputs_embeds: torch.Tensor, attention_mask=None, past_key_values: Optional[KvCache] = None,
                 use_cache: bool = True, return_dict: bool = True, s_max: Optional[int] = None) -> _GptOutput:
        with torch.cuda.device(self._m.device):
            return self._call(inputs_embeds, past_key_values, s_max)

    def _call(self, inputs_embeds, past_key_values, s_max) -> _GptOutput:
        m = self._m
        n_seq, Lnew, H = inputs_embeds.shape
        cache = past_key_values
        if cache is None:
            cache = KvCache(m.dims["gpt_layers"], n_seq, m.dims["gpt_heads"], s_max or min(m.dims["n_pos"], Lnew + 64), 64, m.device)
        emb = inputs_embeds.to(device=m.device, dtype=torch.float32).contiguous()
        ws = m.ws.gpt(n_seq, n_seq * Lnew)
        logits = torch.empty(n_seq, m.dims["vocab_pad"], device=m.device, dtype=torch.float32)
        L.check(L.load().vc_gpt2_forward(C.byref(m.packed.gpt), emb.data_ptr(), n_seq, Lnew, cache.length, C.byref(cache.c),
                                         ws.data_ptr(), ws.numel(), logits.data_ptr(), 0, L.current_stream(m.device)))
        cache.length += Lnew
        # HF returns [B, L, V]; only the last position is ever read on this path (benchmark_baseline.py:210)
        return _GptOutput(logits[:, : m.dims["vocab"]].unsqueeze(1), cache)


class _Decoder:
    def __init__(self, owner: "B200CaptionModel", tokenizer):
        self._m = owner
        self.mapper = _Mapper(owner)
        self.model = _Gpt2(owner)
        self.prefix_len = owner.dims["prefix_len"]
        self.cond_mode = "prefix"
        self.tokenizer = tokenizer

    def _prompt_ids(self, prompt: str) -> List[int]:
        if prompt:
            if self.tokenizer is None:
                raise L.VcError("a tokenizer is needed to encode a non-empty prompt")
            ids = self.tokenizer(prompt, return_tensors="pt").input_ids.reshape(-1).tolist()
            return [int(i) for i in ids]
        return [EOS]                                                      # text_decoder.py:122

    def generate(self, video_emb: torch.Tensor, prompt: str = "", max_new_tokens: int = 32, num_beams: int = 1,
                 temperature: float = 1.0, top_p: float = 0.9, no_repeat_ngram_size: int = 3, repetition_penalty: float = 1.15,
                 min_new_tokens: int = 8) -> List[str]:
        """text_decoder.py:105-146.  Returns decoded strings through `self.tokenizer`; the ids are kept in
        `self.last_ids` / `self.last_lengths` for callers that have no tokenizer (tests, benchmark)."""
        m = self._m
        emb = video_emb.to(device=m.device, dtype=torch.float32)
        B = emb.shape[0]
        prefix = self.mapper(emb).reshape(B, self.prefix_len, m.dims["gpt_dim"])
        do_sample = (num_beams == 1 and temperature != 1.0)               # text_decoder.py:137
        from .decoding import hf_generate_ids
        ids, lengths = hf_generate_ids(m, prefix, self._prompt_ids(prompt), max_new_tokens=max_new_tokens, num_beams=num_beams,
                                       no_repeat_ngram_size=no_repeat_ngram_size, repetition_penalty=repetition_penalty,
                                       min_new_tokens=min_new_tokens, do_sample=do_sample, temperature=temperature, top_p=top_p)
        self.last_ids, self.last_lengths = ids, lengths
        if self.tokenizer is None:
            return ["" for _ in range(B)]
        rows = [ids[b, : int(lengths[b])].tolist() for b in range(B)]
        texts = self.tokenizer.batch_decode(rows, skip_special_tokens=True)
        return [t.strip() for t in texts]


class B200CaptionModel:
    """Returned by `load_caption_model(cfg)` when `cfg.backend == "b200"` (engine.py mirror)."""

    def __init__(self, state_dict: dict, device="cuda:0", *, vit_heads: int = 12, gpt_heads: int = 12, gelu: Optional[str] = None,
                 tokenizer=None, ln_scale: float = 0.6, in_weight: float = 0.4, chunk_frames: int = 256):
        self.device = _require_cuda(device)
        L.load()
        with torch.cuda.device(self.device):
            self.packed: PackedModel = pack(state_dict, self.device, vit_heads=vit_heads, gpt_heads=gpt_heads, gelu=gelu)
        self.dims = self.packed.dims
        self.ln_scale, self.in_weight = float(ln_scale), float(in_weight)
        self.chunk_frames = int(chunk_frames)
        self.ws = Workspace(self)
        self.image_size = int(round((self.dims["tokens"] - 1) ** 0.5)) * self.dims["patch"]     # 224 for B/16 and L/14
        self.resizer = FrameResizer(self.device, self.image_size, self.image_size)
        self.encoder = _Encoder(self)
        self.proj = lambda x: x                                            # nn.Identity (caption_model.py:67)
        self.decoder = _Decoder(self, tokenizer)
        self._pool_buf = None
        self._graphs: dict = {}

    # reference API no-ops so that `load_caption_model(...).to(dev).eval()` style call chains keep working
    def eval(self):
        return self

    def to(self, *_a, **_k):
        return self

    def _pooled(self, B: int) -> torch.Tensor:
        if self._pool_buf is None or self._pool_buf.shape[0] < B:
            self._pool_buf = torch.empty(B, self.dims["vit_dim"], device=self.device, dtype=torch.float32)
        return self._pool_buf[:B]

    # ------------------------------------------------------------------ encode
    def _encode_patches(self, patches: torch.Tensor, n_frames: int) -> torch.Tensor:
        d = self.dims
        chunk = min(self.chunk_frames, n_frames)
        ws = self.ws.vit(chunk)
        cls = self.ws.cls(n_frames)
        L.check(L.load().vc_vit_encode(C.byref(self.packed.vit), patches.data_ptr(), n_frames, chunk, ws.data_ptr(), ws.numel(),
                                       cls.data_ptr(), L.current_stream(self.device)))
        return cls

    def encode_prefix(self, frames_u8: torch.Tensor):
        """uint8 [B,T,H,W,3] on the device -> (feat [B,video_dim] fp32, prefix [B,P,H] fp32).
        Preprocessing + ViT + pool/proj/prefix-norm/mapper; stages named as the reference's NVTX
        ranges (benchmark_baseline.py:93,265,285)."""
        if frames_u8.dtype != torch.uint8 or frames_u8.ndim != 5 or frames_u8.shape[-1] != 3:
            raise ValueError(f"expect uint8 [B,T,H,W,3], got {frames_u8.dtype} {tuple(frames_u8.shape)}")
        if frames_u8.device != self.device:
            raise ValueError("frames must already live on the model's device (use caption_from_host for host buffers)")
        with torch.cuda.device(self.device):
            return self._encode_prefix(frames_u8)

    def _encode_prefix(self, frames_u8: torch.Tensor):
        frames_u8 = frames_u8.contiguous()
        d = self.dims
        if frames_u8.numel() > 0 and tuple(frames_u8.shape[2:4]) != (self.image_size, self.image_size):
            # transforms.Resize((image_size, image_size)) of frame_loader.py:36 — byte-exact with the PIL path, on the device
            frames_u8 = self.resizer(frames_u8)
        B, T, H, W, _ = frames_u8.shape
        n = B * T
        if n == 0:         # empty batch (or no frames): empty outputs, like the reference's modules on a 0-row tensor
            return (torch.zeros(B, d["video_dim"], device=self.device), torch.zeros(B, d["prefix_len"], d["gpt_dim"], device=self.device))
        lib = L.load()
        st = L.current_stream(self.device)
        patches = self.ws.patches(n)
        torch.cuda.nvtx.range_push("Preprocessing")
        L.check(lib.vc_preprocess_u8(frames_u8.data_ptr(), self.packed.lut.data_ptr(), patches.data_ptr(), n, H, W, 1, d["patch"],
                                     d["k_pad"], st))
        torch.cuda.nvtx.range_pop()
        torch.cuda.nvtx.range_push("ViT_Encoder")
        cls = self._encode_patches(patches, n)
        torch.cuda.nvtx.range_pop()
        torch.cuda.nvtx.range_push("Cross_Modal_Alignment")
        feat = torch.empty(B, d["video_dim"], device=self.device, dtype=torch.float32)
        prefix = torch.empty(B, d["prefix_len"], d["gpt_dim"], device=self.device, dtype=torch.float32)
        v = self.packed.vit
        L.check(lib.vc_pool_prefix(cls.data_ptr(), B, T, d["vit_dim"], v.head_w, v.head_b, d["video_dim"], self.ln_scale,
                                   self.in_weight, self.packed.mapper_w.data_ptr(), self.packed.mapper_b.data_ptr(),
                                   d["prefix_len"] * d["gpt_dim"], feat.data_ptr(), prefix.data_ptr(), st))
        torch.cuda.nvtx.range_pop()
        return feat, prefix

    # ------------------------------------------------------------------ decode
    def greedy_ids(self, prefix: torch.Tensor, prompt_ids: Optional[List[int]] = None, max_new_tokens: int = 20,
                   forced_ids: Optional[torch.Tensor] = None, keep_logits: bool = False, use_graph: bool = True):
        """The benchmark's greedy KV-cache loop (benchmark_baseline.py:160-240) with the bookkeeping on the
        device.  Returns (ids int32 [B,max_new] eos-padded, lengths int32 [B], logits or None)."""
        from .decoding import greedy_decode
        with torch.cuda.device(self.device):
            return greedy_decode(self, prefix, prompt_ids or [EOS], max_new_tokens, forced_ids, keep_logits, use_graph)

    def caption_ids(self, frames_u8: torch.Tensor, max_new_tokens: int = 20, num_beams: int = 1, prompt_ids=None, **hf_kwargs):
        """frames -> token ids.  num_beams == 1: benchmark greedy; > 1: HF beam search semantics."""
        feat, prefix = self.encode_prefix(frames_u8)
        if prefix.shape[0] == 0:
            return (torch.full((0, max_new_tokens), EOS, dtype=torch.int32, device=self.device), torch.zeros(0, dtype=torch.int32, device=self.device))
        torch.cuda.nvtx.range_push("GPT2_Decoder_Step")
        try:
            if num_beams == 1 and not hf_kwargs:
                ids, lengths, _ = self.greedy_ids(prefix, prompt_ids, max_new_tokens)
            else:
                from .decoding import hf_generate_ids
                ids, lengths = hf_generate_ids(self, prefix, prompt_ids or [EOS], max_new_tokens=max_new_tokens, num_beams=num_beams,
                                               **hf_kwargs)
        finally:
            torch.cuda.nvtx.range_pop()
        return ids, lengths

    def pipeline(self, max_new_tokens: int = 20, decode_group: int = 2, overlap_decode: bool = True) -> "CaptionPipeline":
        """Throughput path: a three-stream software pipeline over batches (see CaptionPipeline)."""
        return CaptionPipeline(self, max_new_tokens, decode_group, overlap_decode)

    def caption_from_host(self, frames_u8_host: torch.Tensor, max_new_tokens: int = 20, num_beams: int = 1):
        """End-to-end call with HOST buffers: pinned uint8 frames -> H2D -> pipeline -> ids D2H.
        This is the public call bench.py's `e2e` figure times."""
        dev_frames = self.ws.frames(frames_u8_host.shape)
        dev_frames.copy_(frames_u8_host, non_blocking=True)
        ids, lengths = self.caption_ids(dev_frames, max_new_tokens, num_beams)
        return ids.cpu(), lengths.cpu()


def decode_priority() -> int:
    """Stream priority of the decode chain (and of its captured graph's kernel nodes).  Measured on B200 (tools/ab_pipeline.py):
    the same priority as the encoder gives 47.0-47.6 ms per batch, a higher one 49.7 ms - a high-priority chain of ~2000
    tiny kernels keeps interrupting the dispatch of the encoder's large grids.  VC_DECODE_PRIORITY overrides (0 or -1)."""
    return int(os.environ.get("VC_DECODE_PRIORITY", "0"))


class CaptionPipeline:
    """Batches in flight on three streams: H2D copy of batch i+2, preprocess + ViT encode + prefix of batch i+1, and the
    greedy decode of batch i.  The decode step is a chain of ~100 tiny latency-bound kernels that needs a few SM slots,
    the encoder is tensor-pipe bound; with the GEMM's shared-memory footprint cut to 161 KB the decode kernels of the
    previous batches run underneath the encoder kernels of the next one (same stream priority, see decode_priority), so
    steady-state time per batch tends to the encoder time alone.  Results are bit-identical to `caption_ids`: the same
    kernels run on the same data, only on different streams.

        pipe = model.pipeline(max_new_tokens=20)
        t0 = pipe.submit(frames0)      # uint8 [B,T,H,W,3]: device tensor, or pinned host tensor (copied asynchronously)
        t1 = pipe.submit(frames1)
        ids, lens = pipe.result(t0)    # host int32 tensors (pinned); blocks until that batch is done
    """

    MAX_DECODE_ROWS = 512          # sequences decoded as one chain at most (a latency bound: every video of the group waits for the chain;
                                   # 8 x 64 measured 0.7 % faster than 4 x 64 resident and 0.4 % slower from host buffers: bench.py uses 4)

    def __init__(self, model: B200CaptionModel, max_new_tokens: int, decode_group: int = 2, overlap_decode: bool = True):
        """decode_group: consecutive batches whose sequences are decoded together (group * batch <= 256).  The decode
        chain is mostly latency-bound: 20 tokens cost 9.6 ms for 64 sequences, 11.8 ms for 128 and 15.7 ms for 256, so
        decoding several encoder batches per chain cuts the decode cost per batch.  Per-sequence results do not depend
        on the grouping.  overlap_decode=False runs the chain on the encoder's stream (no concurrency)."""
        self.m = model
        self.max_new = int(max_new_tokens)
        self.group = max(1, int(decode_group))
        self.depth = 2 * self.group + 1
        dev = model.device
        with torch.cuda.device(dev):
            self.copy_stream = torch.cuda.Stream(dev)
            self.enc_stream = torch.cuda.Stream(dev)
            self.dec_stream = torch.cuda.Stream(dev, priority=decode_priority()) if overlap_decode else self.enc_stream
        self._slots = [dict(frames=None, ids=None, lens=None, done=None, enc_done=None, h_ids=None, h_lens=None, prefix=None, cb=None, to_host=True)
                       for _ in range(self.depth)]
        self._pending: list = []       # tickets encoded but not yet decoded
        self._n = 0
        self._last_done = None

    def submit(self, frames_u8: torch.Tensor, to_host: bool = True, after_decode=None) -> int:
        m, ticket = self.m, self._n
        slot = self._slots[ticket % self.depth]
        if slot["done"] is not None:
            slot["done"].synchronize()                       # back-pressure: at most `depth` batches in flight
        if frames_u8.device.type == "cpu":
            if slot["frames"] is None or slot["frames"].shape != frames_u8.shape:
                slot["frames"] = torch.empty(frames_u8.shape, dtype=torch.uint8, device=m.device)
            with torch.cuda.stream(self.copy_stream):
                if slot["enc_done"] is not None:
                    self.copy_stream.wait_event(slot["enc_done"])   # the encoder has finished reading this buffer
                slot["frames"].copy_(frames_u8, non_blocking=True)
                copied = self.copy_stream.record_event()
            dev_frames = slot["frames"]
        else:
            # frames already on the device: the encoder stream reads them after whatever the caller's stream has queued, and the
            # caching allocator must not hand the block to somebody else while the encoder may still be reading it
            copied, dev_frames = torch.cuda.current_stream(m.device).record_event(), frames_u8
            frames_u8.record_stream(self.enc_stream)
        with torch.cuda.stream(self.enc_stream):
            if copied is not None:
                self.enc_stream.wait_event(copied)
            feat, prefix = m.encode_prefix(dev_frames)
            slot["enc_done"] = self.enc_stream.record_event()
            prefix.record_stream(self.dec_stream)
        slot["prefix"], slot["cb"], slot["to_host"], slot["done"] = prefix, after_decode, to_host, None
        self._pending.append(ticket)
        self._n += 1
        # decode once the group is complete, or when another batch of this size would not fit into one decode chain
        rows = sum(self._slots[t % self.depth]["prefix"].shape[0] for t in self._pending)
        if len(self._pending) >= self.group or rows + prefix.shape[0] > self.MAX_DECODE_ROWS:
            self._flush()
        return ticket

    def _flush(self) -> None:
        """Decode every encoded-but-undecoded batch as ONE chain."""
        if not self._pending:
            return
        m = self.m
        slots = [self._slots[t % self.depth] for t in self._pending]
        self._pending = []
        with torch.cuda.stream(self.dec_stream):
            for sl in slots:
                self.dec_stream.wait_event(sl["enc_done"])
            prefix = slots[0]["prefix"] if len(slots) == 1 else torch.cat([sl["prefix"] for sl in slots], dim=0)
            ids, lens, _ = m.greedy_ids(prefix, None, self.max_new)
            o = 0
            for sl in slots:
                B = sl["prefix"].shape[0]
                if sl["ids"] is None or sl["ids"].shape != (B, ids.shape[1]):
                    sl["ids"] = torch.empty(B, ids.shape[1], dtype=ids.dtype, device=m.device)
                    sl["lens"] = torch.empty(B, dtype=lens.dtype, device=m.device)
                    sl["h_ids"] = torch.empty(sl["ids"].shape, dtype=ids.dtype).pin_memory()
                    sl["h_lens"] = torch.empty(sl["lens"].shape, dtype=lens.dtype).pin_memory()
                sl["ids"].copy_(ids[o:o + B]); sl["lens"].copy_(lens[o:o + B])   # the decode graph's static outputs are reused
                o += B
                if sl["cb"] is not None:
                    sl["cb"](sl["ids"], sl["lens"])            # e.g. the NCCL id gather, enqueued on the decode stream
                if sl["to_host"]:
                    sl["h_ids"].copy_(sl["ids"], non_blocking=True)
                    sl["h_lens"].copy_(sl["lens"], non_blocking=True)
                sl["prefix"] = None
            done = self.dec_stream.record_event()
            for sl in slots:
                sl["done"] = done
            self._last_done = done

    def result(self, ticket: int, host: bool = True):
        if ticket < self._n - self.depth or ticket >= self._n:
            raise L.VcError(f"ticket {ticket} is no longer (or not yet) in flight")
        if ticket in self._pending:
            self._flush()
        slot = self._slots[ticket % self.depth]
        slot["done"].synchronize()
        return (slot["h_ids"], slot["h_lens"]) if host else (slot["ids"], slot["lens"])

    def last_event(self):
        self._flush()
        return self._last_done

    def warm(self, frames_u8: torch.Tensor, after_decode=None) -> None:
        """Run every decode-group size once (1..decode_group batches of this shape) and touch every pipeline slot, so that
        all CUDA graphs, KV caches, workspaces, per-slot device frame buffers and pinned result buffers the steady state
        will need exist before the first timed / latency-sensitive batch (a cudaMalloc or a pinned allocation inside the
        stream of batches synchronises the device).  Pass a host tensor to warm the host-buffer path."""
        to_host = frames_u8.device.type == "cpu"
        for g in range(1, self.group + 1):
            for _ in range(g):
                self.submit(frames_u8, to_host=to_host, after_decode=after_decode)
            self.drain()
        for _ in range(self.depth):
            self.submit(frames_u8, to_host=to_host, after_decode=after_decode)
        self.drain()

    def drain(self) -> None:
        self._flush()
        for s in (self.copy_stream, self.enc_stream, self.dec_stream):
            s.synchronize()
