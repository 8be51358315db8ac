import sys
sys.path.insert(0, "/root/repo")
import torch, vcb200
from vcb200 import synthetic
from vcb200.model import B200CaptionModel
a = synthetic.ARCHS["tiny"]
m = B200CaptionModel(synthetic.make_state_dict(a, seed=1234), "cuda:0", vit_heads=a.vit_heads, gpt_heads=a.gpt_heads, chunk_frames=8)
for B, T in [(1, 1), (3, 5), (0, 4), (2, 16), (5, 3)]:
    f = synthetic.make_batch_u8(0, B, T).cuda() if B > 0 else torch.zeros(0, T, 224, 224, 3, dtype=torch.uint8, device="cuda")
    try:
        ids, lens = m.caption_ids(f, max_new_tokens=4)
        torch.cuda.synchronize()
        print(B, T, "ok", tuple(ids.shape), lens.tolist())
    except Exception as e:
        print(B, T, "ERR", type(e).__name__, str(e)[:200])
# max_new_tokens = 1, long decode near the cache limit
f = synthetic.make_batch_u8(0, 2, 2).cuda()
for n in (1, 2, 64):
    try:
        ids, lens = m.caption_ids(f, max_new_tokens=n); torch.cuda.synchronize(); print("max_new", n, "ok", tuple(ids.shape))
    except Exception as e:
        print("max_new", n, "ERR", type(e).__name__, str(e)[:200])
# beam with B=1
try:
    ids, lens = m.caption_ids(f[:1], max_new_tokens=6, num_beams=3); torch.cuda.synchronize(); print("beam B=1 ok", ids.tolist(), lens.tolist())
except Exception as e:
    print("beam ERR", type(e).__name__, str(e)[:200])
