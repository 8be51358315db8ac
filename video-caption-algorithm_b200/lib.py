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
SOURCES = ["gemm_tcgen05.cu", "vit_kernels.cu", "resize_kernels.cu", "vit_attention_tc.cu", "gpt2_kernels.cu", "skinny_gemm.cu", "decode_chain.cu", "beam_kernels.cu", "c_abi.cu"]
HEADERS = ["vc_common.cuh", "vc_kernels.h", "../../include/vcb200.h"]
# -cudart shared: the runtime is the process's libcudart.so (torch ships one), not a private static copy inside the library
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "shared"]
ABI_VERSION = 5          # == vc_abi_version(); bumped with every struct / signature change (include/vcb200.h)


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
    """nvcc -gencode arch=compute_100a,code=sm_100a … -> csrc/libvcb200.so (cross-compiles without a GPU).
    One object per translation unit (only the stale ones are recompiled, in parallel), then one link."""
    if not force and not needs_build():
        return SO_PATH
    from concurrent.futures import ThreadPoolExecutor
    obj_dir = CSRC / "build"
    obj_dir.mkdir(exist_ok=True)
    nvcc = _nvcc()
    hdr_t = max((CSRC / h).stat().st_mtime for h in HEADERS if (CSRC / h).exists())
    compile_flags = [f for f in NVCC_FLAGS if f not in ("-shared",)]
    logs = []

    def one(src: str):
        obj = obj_dir / (src[:-3] + ".o")
        sp = CSRC / src
        if not force and obj.exists() and obj.stat().st_mtime > max(sp.stat().st_mtime, hdr_t):
            return obj, None
        cmd = [nvcc, *compile_flags, "-c", "-o", str(obj), str(sp)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            return obj, VcError(f"nvcc failed on {src}:\n" + r.stdout + r.stderr)
        logs.append(r.stderr)
        return obj, None

    srcs = [f for f in SOURCES if (CSRC / f).exists()]
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(one, srcs))
    for _, err in results:
        if err is not None:
            raise err
    r = subprocess.run([nvcc, *NVCC_FLAGS, "-o", str(SO_PATH), *[str(o) for o, _ in results]], capture_output=True, text=True)
    if r.returncode != 0:
        raise VcError("nvcc link failed:\n" + r.stdout + r.stderr)
    if verbose:
        print("".join(logs))
    return SO_PATH


# ------------------------------------------------------------------ C structs (include/vcb200.h)
_p = C.c_void_p


class VcVitLayer(C.Structure):
    _fields_ = [(n, _p) for n in ("ln1_g", "ln1_b", "qkv_w", "qkv_b", "proj_w", "proj_b", "ln2_g", "ln2_b",
                                  "fc1_w", "fc1_b", "fc2_w", "fc2_b",
                                  "qkv_wf", "qkv_cs", "qkv_bf", "fc1_wf", "fc1_cs", "fc1_bf")]


class VcVitWeights(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dim", "layers", "heads", "mlp", "tokens", "patch_k", "gelu_tanh", "video_dim")] + \
               [(n, _p) for n in ("patch_w", "patch_b", "cls_pos0", "pos", "lnf_g", "lnf_b", "head_w", "head_b")] + \
               [("layer", C.POINTER(VcVitLayer))]


class VcGptLayer(C.Structure):
    _fields_ = [(n, _p) for n in ("ln1_g", "ln1_b", "attn_w", "attn_b", "aproj_w", "aproj_b", "ln2_g", "ln2_b",
                                  "fc_w", "fc_b", "mproj_w", "mproj_b",
                                  "attn_wf", "attn_cs", "attn_bf", "fc_wf", "fc_cs", "fc_bf")]


class VcGptWeights(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("dim", "layers", "heads", "vocab", "n_pos", "vocab_pad")] + \
               [(n, _p) for n in ("wte", "wpe", "lnf_g", "lnf_b", "lmh_w", "lmh_cs", "lmh_b")] + [("layer", C.POINTER(VcGptLayer))]


class VcBeamState(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("B", "nb", "max_len", "eos")] + \
               [(n, _p) for n in ("running_scores", "running_seqs", "fin_seqs", "fin_scores", "fin_done", "fin_len", "unsatisfied",
                                  "flags", "stopped", "src_rows", "next_tok")]


class VcKvCache(C.Structure):
    _fields_ = [("kv", _p), ("slot", _p)] + [(n, C.c_int32) for n in ("layers", "n_seq", "heads", "s_max", "head_dim")]


_i, _f, _sz = C.c_int, C.c_float, C.c_size_t
_SIGNATURES = {
    "vc_last_error": (C.c_char_p, []),
    "vc_abi_version": (_i, []),
    "vc_num_sms": (_i, []),
    "vc_launch_count": (C.c_longlong, []),
    "vc_debug_trace": (_i, [_p, _i]),
    "vc_prof_begin": (_i, []),
    "vc_prof_end": (_i, [_i, _p, _p, _p, _p]),
    "vc_resize_bilinear_u8": (_i, [_p, _i, _i, _i, _p, _p, _i, _i, _p, _p, _i, _p, _p, _i, _p]),
    "vc_preprocess_u8": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "vc_patchify_f32": (_i, [_p, _p, _i, _i, _i, _i, _i, _p]),
    "vc_gemm_bf16": (_i, [_p, _p, _p, _i, _i, _i, _i, _p, _i, _p, _i, _p]),
    "vc_gemm_resid_stats": (_i, [_p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _f, _i, _p]),
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
    "vc_beam_init": (_i, [C.POINTER(VcBeamState), _p]),
    "vc_beam_update": (_i, [C.POINTER(VcBeamState), _p, _p, _i, _i, _f, _p]),
    "vc_beam_finalize": (_i, [C.POINTER(VcBeamState), _p, _p, _p]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """dlopen the in-tree library (RTLD_GLOBAL not needed: plain C ABI)."""
    global _lib, SO_PATH
    if _lib is not None:
        return _lib
    if os.environ.get("VC_LIB"):                 # A/B builds of the same ABI (tools only): load this file, never rebuild it
        SO_PATH = Path(os.environ["VC_LIB"])
        build_if_missing = False
    if build_if_missing and needs_build():
        # on the GPU box the prebuilt .so travels with the snapshot; rebuild when stale.  A failed rebuild is an error:
        # binding today's signatures to yesterday's binary would run outdated kernels with mismatched argument lists.
        build()
    if not SO_PATH.exists():
        raise VcError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback)")
    lib = C.CDLL(str(SO_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here = header/library drift
        fn.restype = res
        fn.argtypes = args
    got = lib.vc_abi_version()
    if got != ABI_VERSION:
        raise VcError(f"{SO_PATH} has ABI version {got}, the Python binding expects {ABI_VERSION}: rebuild the library")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().vc_last_error()
        raise VcError(f"libvcb200 error {status}: {msg.decode() if msg else '?'}")


def ptr(t) -> int:
    """device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def current_stream(device=None) -> int:
    """torch's current stream ON `device` (not on whatever device happens to be current)."""
    import torch
    return torch.cuda.current_stream(device).cuda_stream
