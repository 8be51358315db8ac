"""Soak: several hundred batches through CaptionPipeline from pinned host buffers; device memory must stay flat and every
batch of the same frames must give the same ids."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
a = synthetic.ARCHS["vit_b16_gpt2"]
m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, chunk_frames=1024)
host = [synthetic.make_batch_u8(64 * k, 64, 16).pin_memory() for k in range(2)]
pipe = m.pipeline(max_new_tokens=20, decode_group=4)
pipe.warm(host[0])
torch.cuda.synchronize()
mem0 = torch.cuda.memory_allocated()
ref = {}
tickets = []
bad = 0
for i in range(n):
    tickets.append((pipe.submit(host[i % 2]), i % 2))
    if len(tickets) > 6:
        t, k = tickets.pop(0)
        ids, lens = pipe.result(t)
        key = (ids.clone(), lens.clone())
        if k not in ref:
            ref[k] = key
        elif not (torch.equal(ref[k][0], key[0]) and torch.equal(ref[k][1], key[1])):
            bad += 1
for t, k in tickets:
    ids, lens = pipe.result(t)
    if not (torch.equal(ref[k][0], ids) and torch.equal(ref[k][1], lens)):
        bad += 1
pipe.drain()
mem1 = torch.cuda.memory_allocated()
print(f"{n} batches: mismatching batches {bad}; device memory {mem0 / 2**20:.0f} MiB -> {mem1 / 2**20:.0f} MiB; reserved {torch.cuda.memory_reserved() / 2**20:.0f} MiB")
assert bad == 0 and mem1 <= mem0 + (64 << 20)
