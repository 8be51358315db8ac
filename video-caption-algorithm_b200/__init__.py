"""B200-native video-caption inference hot path (ViT encode -> visual prefix ->
GPT-2 decode) behind the reference's `model.encoder / model.decoder` surface.

Device work is hand-written sm_100a CUDA in `csrc/`, reached through the C-ABI
declared in `include/vcb200.h` (built to `csrc/libvcb200.so`).  There is no CPU
or PyTorch fallback: without the library or a GPU the compute entry points raise.
"""
from . import synthetic  # noqa: F401

__all__ = ["synthetic"]
