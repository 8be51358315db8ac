"""How far inside the stated tolerances the CUDA path sits (the tests only assert): features and teacher-forced logits against the
reference's own fp32 outputs (golden fixtures) and against the oracle at BASELINE sizes.  Writes profiles/r2_parity_margins.json."""
import json
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import vcb200  # noqa
from vcb200 import synthetic
from vcb200.model import B200CaptionModel
from oracle import vc_oracle as O

DEV = "cuda:0"
out = {"tolerances": {"feature_maxabs": 0.02, "feature_cos": 0.9999, "logit_maxabs": 0.06, "logit_cos": 0.9995, "tf_agreement": 0.93}}


def cosmin(a, b):
    return torch.nn.functional.cosine_similarity(a.flatten(1).float(), b.flatten(1).float(), dim=-1).min().item()


for arch in ("tiny", "vit_b16_gpt2"):
    g = np.load(ROOT / "tests" / "golden" / f"path_{arch}.npz")
    a = synthetic.ARCHS[arch]
    m = B200CaptionModel(synthetic.make_state_dict(a, seed=int(g["seed"])), DEV, vit_heads=a.vit_heads, gpt_heads=a.gpt_heads)
    B, T, n_new = int(g["B"]), int(g["T"]), int(g["max_new_tokens"])
    feat, prefix = m.encode_prefix(synthetic.make_batch_u8(0, B, T).to(DEV))
    ref_feat, ref_prefix, ref_ids = torch.from_numpy(g["feat"]), torch.from_numpy(g["prefix"]), torch.from_numpy(g["ids"])
    _, _, logits = m.greedy_ids(ref_prefix.to(DEV), None, n_new, forced_ids=ref_ids.to(DEV), keep_logits=True)
    torch.cuda.synchronize()
    steps, stride = g["logits_sub"].shape[0], int(g["logits_stride"])
    lg, ref_sub = logits[:steps].cpu(), torch.from_numpy(g["logits_sub"])
    top = torch.from_numpy(g["logits_top_idx"])[:, :, 0]
    out[f"reference_golden_{arch}"] = {
        "what": f"{B} videos x {T} frames, {n_new} tokens, vs the unmodified reference modules in fp32",
        "feature_maxabs": (feat.cpu() - ref_feat).abs().max().item(), "feature_cos_min": cosmin(feat.cpu(), ref_feat),
        "logit_maxabs": (lg[:, :, ::stride] - ref_sub).abs().max().item(),
        "logit_cos_min": cosmin(lg[:, :, ::stride].reshape(steps * B, -1), ref_sub.reshape(steps * B, -1)),
        "tf_agreement": (lg.argmax(-1) == top).float().mean().item(),
        "reference_own_bf16_feature_maxabs": float(g["stat_feat_bf16_maxabs"]), "reference_own_bf16_logit_maxabs": float(g["stat_logits_bf16_maxabs"]),
        "reference_own_bf16_tf_agreement": float(g["stat_tf_agree_bf16"])}
    if arch == "vit_b16_gpt2":
        # cfg2 size against the oracle: 64 videos x 16 frames
        sd = synthetic.make_state_dict(a, seed=int(g["seed"]))
        fr = synthetic.make_batch_u8(0, 64, 16)
        f64, p64 = m.encode_prefix(fr.to(DEV))
        torch.cuda.synchronize()
        fo = torch.cat([O.encode(sd, O.preprocess_u8(fr[i:i + 8]), a.vit_heads) for i in range(0, 64, 8)], 0)
        po = O.visual_prefix(sd, fo)
        ids_o, _, lg_o = O.greedy_decode(sd, po[:8], torch.tensor([[50256]]), 20, heads=a.gpt_heads, keep_logits=True)
        Lo = torch.stack(lg_o, 0)
        forced = torch.full((64, 20), 11, dtype=torch.int32); forced[:8] = ids_o.int()
        pre = p64.clone(); pre[:8] = po[:8].to(DEV)
        _, _, lg64 = m.greedy_ids(pre, None, 20, forced_ids=forced.to(DEV), keep_logits=True)
        torch.cuda.synchronize()
        l8 = lg64[: Lo.shape[0], :8].cpu()
        out["oracle_cfg2_64x16"] = {"what": "64 videos x 16 frames full depth vs the oracle (features: all 64 videos; logits: 8 videos x 20 teacher-forced steps, 64-row chain)",
                                    "feature_maxabs": (f64.cpu() - fo).abs().max().item(), "feature_cos_min": cosmin(f64.cpu(), fo),
                                    "logit_maxabs": (l8 - Lo).abs().max().item(), "logit_cos_min": cosmin(l8.flatten(0, 1), Lo.flatten(0, 1)),
                                    "tf_agreement": (l8.argmax(-1) == Lo.argmax(-1)).float().mean().item()}
    del m
    torch.cuda.empty_cache()
(ROOT / "profiles" / "r2_parity_margins.json").write_text(json.dumps(out, indent=1))
print(json.dumps(out, indent=1))
