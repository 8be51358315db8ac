"""Timings of the other BASELINE.json configs (parity-test cases, not bench lines): cfg4 beam-5 / 30 tokens at batch 64,
cfg5 ViT-L/14 + GPT-2 medium, 32 frames, 16 videos per GPU (the 8-GPU share of batch 128)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel


def timeit(fn, n=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "cfg4"):
    a = synthetic.ARCHS["vit_b16_gpt2"]
    m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, chunk_frames=1024)
    frames = synthetic.make_batch_u8(0, 64, 16).cuda()
    feat, prefix = m.encode_prefix(frames)
    ms = timeit(lambda: m.caption_ids(frames, max_new_tokens=30, num_beams=5, no_repeat_ngram_size=3, repetition_penalty=1.1), n=2)
    enc = timeit(lambda: m.encode_prefix(frames), n=3)
    print(f"cfg4  ViT-B/16 + GPT-2 small, 64 videos x 16 frames, beam 5 x 30 tokens: {ms:.1f} ms per batch ({64 / ms * 1e3:.0f} captions/s), encode {enc:.1f} ms")
    del m
    torch.cuda.empty_cache()
if which in ("all", "cfg5"):
    a = synthetic.ARCHS["vit_l14_gpt2m"]
    m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, chunk_frames=512)
    frames = synthetic.make_batch_u8(0, 16, 32).cuda()
    enc = timeit(lambda: m.encode_prefix(frames), n=3)
    tot = timeit(lambda: m.caption_ids(frames, max_new_tokens=20), n=3)
    gflop = 162.024 * 16 * 32
    print(f"cfg5  ViT-L/14 + GPT-2 medium, 16 videos x 32 frames (1/8 of batch 128), greedy 20: {tot:.1f} ms per batch ({16 / tot * 1e3:.0f} captions/s per GPU), "
          f"encode {enc:.1f} ms = {gflop / enc:.0f} TFLOP/s")
