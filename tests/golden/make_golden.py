"""Generates the committed golden vectors from the REFERENCE's own source files (executed by path through
oracle/ref_shim.py).  Run here (where /root/reference exists):   python tests/golden/make_golden.py
Weights are key-seeded (tests/model_factory.py), so only inputs + reference outputs are stored."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from tests import model_factory as mf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (swin, T, H, W, Q, model kwargs, targets builder)
    "det_noprompt_ws4": (mf.TINY_SWIN, 2, 64, 96, 10, dict(enc_layers=2, dec_layers=4),
                         dict(task="detection", dataset_name="ytvis21", prompt_type="visual", frame_indices=[3, 4])),
    "det_noprompt_ws7": (mf.SMALL_SWIN7, 3, 96, 160, 12, dict(enc_layers=2, dec_layers=3),
                         dict(task="detection", dataset_name="ovis", prompt_type="visual", frame_indices=[0, 1, 2])),
    "det_category_prompts": (mf.TINY_SWIN, 2, 64, 64, 6, dict(enc_layers=1, dec_layers=3),
                             dict(task="detection", dataset_name="bdd_track", prompt_type="text", frame_indices=[0, 1])),
    "grounding_l2v": (mf.TINY_SWIN, 2, 64, 64, 6,
                      dict(enc_layers=1, dec_layers=3, text_prompt_to_image_enable=True, self_attn_mask_type="sep-blocked"),
                      dict(task="grounding", dataset_name="refytvos", prompt_type="text", frame_indices=[0, 1], num_exp=3)),
}


def build_targets(spec, T, seed):
    tg = {k: v for k, v in spec.items() if k not in ("frame_indices", "num_exp")}
    tg["frame_indices"] = torch.tensor(spec["frame_indices"])
    if spec["task"] == "grounding":
        g = torch.Generator().manual_seed(seed + 77)
        P = spec["num_exp"]
        tg["exp_word_feats"] = torch.randn(P, 77, T, 640, generator=g)
        tg["exp_sentence_feats"] = torch.randn(P, T, 640, generator=g)
        tg["exp_word_len"] = torch.full((P,), 10)
    return [tg]


def main():
    ref = ref_shim.load()
    for name, (swin, T, H, W, Q, kw, tspec) in CASES.items():
        clip = mf.make_clip_emb()
        bb, pix, dec = ref_shim.build_reference_model(swin, num_queries=Q, num_frames=T, clip_emb=clip, **kw)
        for m in (bb, pix, dec):
            m.load_state_dict(mf.keyed_state_dict(m.state_dict()))
        g = torch.Generator().manual_seed(42)
        frames = torch.randn(T, 3, H, W, generator=g)
        feats, (mfeat, ms), out = ref_shim.reference_clip_forward(bb, pix, dec, frames, build_targets(tspec, T, 0))
        blob = {
            "case": name, "frames": frames,
            "res5": feats["res5"].clone(), "res2_mean": feats["res2"].mean((2, 3)),
            "mask_features_c8": mfeat[:, ::8].clone(), "ms0": ms[0].clone(),
            "pred_masks": out["pred_masks"].clone(), "pred_logits": out["pred_logits"].clone(),
            "pred_embds": out["pred_embds"].clone(),
            "pred_reid_logits": out["pred_reid_logits"] if torch.is_tensor(out["pred_reid_logits"]) else None,
        }
        torch.save(blob, os.path.join(HERE, f"{name}.pt"))
        print(name, {k: tuple(v.shape) for k, v in blob.items() if torch.is_tensor(v)})
    # operator-level: the reference's own MSDeformAttn test inputs (ops/test.py:24-41), fp32
    torch.manual_seed(3)
    shapes = [(6, 4), (3, 2)]
    value = torch.rand(1, 30, 2, 2) * 0.01
    loc = torch.rand(1, 2, 2, 2, 2, 2)
    w = torch.rand(1, 2, 2, 2, 2) + 1e-5
    w = w / w.sum(-1, keepdim=True).sum(-2, keepdim=True)
    out = ref.ms_deform_attn_core_pytorch(value, shapes, loc, w)
    torch.save({"shapes": shapes, "value": value, "loc": loc, "w": w, "out": out}, os.path.join(HERE, "msda_ops_test.pt"))


if __name__ == "__main__":
    main()
