"""ctypes binding of `csrc/libvcb200.so` (C ABI: include/vcb200.h) and its build recipe.

The library is built in-tree with nvcc for sm_100a only.  There is no fallback:
`load()` raises if the shared object is missing and cannot be built, and every
wrapper raises `VcError` on a non-zero status (the reference's CuPy hooks return
None / fall back to torch instead — core/operators/cupy_vit_pool.py:185-186,
cupy_linear_mapper.py:175-184 — which north_star forbids here).
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
SO_PATH = CSRC / "libvcb200.so"
SOURCES = ["gemm_tcgen05.cu", "vit_kernels.cu", "resize_kernels.cu", "vit_attention_tc.cu", "gpt2_kernels.cu", "skinny_gemm.cu", "decode_step.cu", "decode_lean.cu", "beam_kernels.cu", "c_abi.cu"]
HEADERS = ["vc_common.cuh", "vc_kernels.h", "../../include/vcb200.h"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


class VcError(RuntimeError):
    pass


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise VcError("nvcc not found: cannot build libvcb200.so")
    return exe


def needs_build() -> bool:
    if not SO_PATH.exists():
        return True
    t = SO_PATH.stat().st_mtime
    return any((CSRC / f).exists() and (CSRC / f).stat().st_mtime > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> Path:
    """nvcc -gencode arch=compute_100a,code=sm_100a … -> csrc/libvcb200.so (cross-compiles without a GPU)."""
    if not force and not needs_build():
        return SO_PATH
    srcs = [str(CSRC / f) for f in SOURCES if (CSRC / f).exists()]
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(SO_PATH), *srcs]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise VcError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stderr)
    return SO_PATH


# ------------------------------------------------------------------ C structs (include/vcb200.h)
_p = C.c_void_p


class VcVitLayer(C.Structure):
    _fields_ = [(n, _p) for n in ("ln1_g", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_g", "ln2_b",
                                  "fc1_w", "fc1_b", "fc2_w", "fc2_b")]


class VcVitWeights(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dim", "layers", "heads", "mlp", "tokens", "patch_k", "gelu_tanh", "video_dim")] + \
               [(n, _p) for n in ("patch_w", "patch_b", "cls_pos0", "pos", "lnf_g", "lnf_b", "head_w", "head_b")] + \
               [("layer", C.POINTER(VcVitLayer))]


class VcGptLayer(C.Structure):
    _fields_ = [(n, _p) for n in ("ln1_g", "ln1_b", "attn_w", "attn_b", "aproj_w", "aproj_b", "ln2_g", "ln2_b",
                                  "fc_w", "fc_b", "mproj_w", "mproj_b")]


class VcGptWeights(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dim", "layers", "heads", "vocab", "n_pos", "vocab_pad")] + \
               [(n, _p) for n in ("wte", "wpe", "lnf_g", "lnf_b")] + [("layer", C.POINTER(VcGptLayer))]


class VcKvCache(C.Structure):
    _fields_ = [("kv", _p), ("slot", _p)] + [(n, C.c_int32) for n in ("layers", "n_seq", "heads", "s_max", "head_dim")]


_i, _f, _sz = C.c_int, C.c_float, C.c_size_t
_SIGNATURES = {
    "vc_last_error": (C.c_char_p, []),
    "vc_abi_version": (_i, []),
    "vc_num_sms": (_i, []),
    "vc_launch_count": (C.c_longlong, []),
    "vc_prof_begin": (_i, []),
    "vc_prof_end": (_i, [_i, _p, _p, _p, _p]),
    "vc_resize_bilinear_u8": (_i, [_p, _i, _i, _i, _p, _p, _i, _i, _p, _p, _i, _p, _p, _i, _p]),
    "vc_preprocess_u8": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "vc_patchify_f32": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "vc_gemm_bf16": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _i, _p, _i, _p]),
    "vc_layernorm_f32_bf16": (_i, [_p, _p, _p, _p, _i, _i, _f, _p]),
    "vc_vit_attention": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "vc_vit_attention_mma_sync": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "vc_vit_workspace_bytes": (_sz, [C.POINTER(VcVitWeights), _i]),
    "vc_vit_encode": (_i, [C.POINTER(VcVitWeights), _p, _i, _i, _p, _sz, _p, _p]),
    "vc_pool_prefix": (_i, [_p, _i, _i, _i, _p, _p, _i, _f, _f, _p, _p, _i, _p, _p, _p]),
    "vc_vit_pool_temporal": (_i, [_p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "vc_linear_bias_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "vc_gpt_workspace_bytes": (_sz, [C.POINTER(VcGptWeights), _i, _i]),
    "vc_gpt2_forward": (_i, [C.POINTER(VcGptWeights), _p, _i, _i, _i, C.POINTER(VcKvCache), _p, _sz, _p, _p, _p]),
    "vc_gpt2_embed_tokens": (_i, [C.POINTER(VcGptWeights), _p, _i, _p, _p]),
    "vc_greedy_decode": (_i, [C.POINTER(VcGptWeights), _p, _i, _i, _p, _i, _i, _i, C.POINTER(VcKvCache), _p, _sz, _p, _p,
                              _p, _p, _p]),
    "vc_argmax_f32": (_i, [_p, _i, _i, _p, _p]),
    "vc_skinny_ksplit": (_i, [_i, _i]),
    "vc_skinny_gemm_partial": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "vc_beam_step": (_i, [_p, C.c_longlong, _i, _i, _i, _p, _i, _i, _p, _f, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p]),
    "vc_beam_reorder": (_i, [_p, _p, _p, _i, _i, _i, _p]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the in-tree library (RTLD_GLOBAL not needed: plain C ABI)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and needs_build():
        # on the GPU box the prebuilt .so travels with the snapshot; rebuild only when stale and nvcc exists
        try:
            build()
        except VcError:
            if not SO_PATH.exists():
                raise
    if not SO_PATH.exists():
        raise VcError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback)")
    lib = C.CDLL(str(SO_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library drift
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().vc_last_error()
        raise VcError(f"libvcb200 error {status}: {msg.decode() if msg else '?'}")


def ptr(t) -> int:
    """device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
