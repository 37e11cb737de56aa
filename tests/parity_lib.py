"""Full-geometry parity of the CUDA path against the CPU oracle for the BASELINE.json configurations (SURVEY.md 8d
"Synthetic inputs"), shared by tests/test_parity_full_geometry.py (the `-m gpu` tests) and tests/tools/parity_*.py.

  ns  Swin-L, 720x1280 (pads to 736x1280), Q=200, T=5, detection, no prompts          (north-star metric)
  c2  Swin-T, 480x864, Q=100, T=5, detection (VIS inference path)
  c3  Swin-B, 720x1280, Q=200, task "sot": P=10 objects given as first-frame rectangle masks (area 2-20 % of the frame,
      seed 1), R=128 points per prompt, prompt memory grown over 3 consecutive stride-1 clips of T=5
  c4  Swin-L, 720x1280, Q=200 + P=32 text prompts, task "grounding", self-attention mask "sep-blocked", lang->vision on, T=10
  c5  Swin-L, 1080x1920 (pads to 1088x1920), Q=200, T=8, detection (single-GPU result at the sharded config's geometry)

The decoder takes hard decisions: an attention-mask bit is the sign of a mask logit, and a query whose mask row is
empty attends everywhere (..._univs.py:390, 555-566).  A logit within fp32 rounding distance of zero (|m| ~ 1e-7 of the
logit scale -- the oracle's own sign for it is arbitrary) flips a decision and moves THAT query by far more than any
arithmetic error.  So every clip is run three times and three things are reported:
  oracle          CPU, records the attention-mask bits of every prediction head
  free-running    CUDA path on its own decisions: errors, number of flipped bits, which queries own them
  same-decisions  CUDA path replaying the oracle's bits: isolates the arithmetic error (tolerance 1e-3, north star)
Tolerance metric: max|a-b| / max|b| per tensor."""
import copy
import os
import time

import torch

from oracle.cpu_backend import oracle_ops
from univs_b200.build import build_model, make_cfg
from univs_b200.precision import get_precision, set_precision

TOL = 1e-3
CONFIGS = {
    # name: (variant, H, W, Q, T, task, clips)
    "ns": ("large", 720, 1280, 200, 5, "detection", 1),
    "c2": ("tiny", 480, 864, 100, 5, "detection", 1),
    "c3": ("base", 720, 1280, 200, 5, "sot", 3),
    "c4": ("large", 720, 1280, 200, 10, "grounding", 1),
    "c5": ("large", 1080, 1920, 200, 8, "detection", 1),
}


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def _popcount32(x):
    x = x.to(torch.int64) & 0xFFFFFFFF
    x = x - ((x >> 1) & 0x55555555)
    x = (x & 0x33333333) + ((x >> 2) & 0x33333333)
    x = (x + (x >> 4)) & 0x0F0F0F0F
    return (x * 0x01010101 >> 24) & 0xFF


from univs_b200.synthetic import ClipSource, clone_targets as _clone_targets, rectangle_masks, univs_overrides  # noqa: E402,F401


def run_config(name, T=None, clips=None, precision="fp16x3", variant=None, hw=None, points=128, threads=None):
    """Runs configuration `name` (CONFIGS) on the CPU oracle and, when a GPU is present, on the CUDA path; returns the
    report dict (see module docstring).  T / clips / variant / hw override the BASELINE values (quick plumbing checks)."""
    v0, H, W, Q, T0, task, clips0 = CONFIGS[name]
    variant, T = variant or v0, int(T or T0)
    clips = int(clips or clips0) if task == "sot" else 1
    if hw is not None:
        H, W = hw
    dry = not torch.cuda.is_available()        # no GPU: oracle side only (checks the target construction)
    V = T + clips - 1
    src = ClipSource(task, T, V, H, W)
    cfg = make_cfg(variant, Q, T, clip_emb=src.clip_emb, **univs_overrides(task, points))
    cpu_model = build_model(cfg)
    gpu_model = None
    if not dry:
        gpu_model = build_model(cfg).cuda()
        gpu_model.load_state_dict(cpu_model.state_dict())
    old_threads = torch.get_num_threads()
    torch.set_num_threads(threads or min(32, os.cpu_count() or 1))
    P = src.P
    targets, clip_inputs = src.targets, src.clip_inputs

    res = {"config": name, "geometry": f"Swin-{variant} T={T} {H}x{W} Q={Q} P={P} task={task} clips={clips}",
           "precision": precision, "tolerance": TOL, "tolerance_metric": "max|a-b|/max|b| per tensor", "clips": []}
    ctg, gtg = targets("cpu"), (None if dry else targets("cuda"))
    saved_precision = get_precision()
    set_precision(precision)
    try:
        for c in range(clips):
            recorded = {}
            cdec = cpu_model.sem_seg_head.predictor
            cdec.attn_mask_hook = lambda i, b, r: (recorded.__setitem__(i, (b.clone(), r.clone())) or (b, r))
            t0 = time.time()
            torch.manual_seed(100 + c)
            with oracle_ops():
                wout = cpu_model.clip_forward(clip_inputs(ctg, c, "cpu"), ctg)
            cdec.attn_mask_hook = None
            cpu_s = time.time() - t0
            row = {"clip": c, "cpu_oracle_seconds": round(cpu_s, 2), "shape": list(wout["pred_masks"].shape)}
            if dry:
                res["clips"].append(row)
                continue
            gdec = gpu_model.sem_seg_head.predictor
            nq = wout["pred_masks"].shape[1]
            flips_q = torch.zeros(nq, dtype=torch.int64)
            total_bits = [0]

            def count(i, b, r):
                ob = recorded[i][0]
                flips_q.add_(_popcount32(b.cpu() ^ ob).sum(-1).sum(0))        # bits [T,Q,words] -> per query
                total_bits[0] += ob.numel() * 32
                return b, r
            before = _clone_targets(gtg)
            x = clip_inputs(gtg, c, "cuda")
            gdec.attn_mask_hook = count
            torch.manual_seed(100 + c)
            gout = gpu_model.clip_forward(x, gtg)
            # replay of the oracle's decisions from the same pre-clip state (prompt memory pool)
            rtg = before
            x = clip_inputs(rtg, c, "cuda")
            gdec.attn_mask_hook = lambda i, b, r: (recorded[i][0].cuda(), recorded[i][1].cuda())
            torch.manual_seed(100 + c)
            fout = gpu_model.clip_forward(x, rtg)
            gdec.attn_mask_hook = None
            torch.cuda.synchronize()
            pm, wm = gout["pred_masks"].float().cpu(), wout["pred_masks"]
            perq = (pm - wm).abs().flatten(2).amax(2)[0] / wm.abs().max()
            beyond = perq > TOL
            flipped = flips_q > 0
            row.update({
                "free_running": {
                    "pred_masks": rel(pm, wm), "pred_logits": rel(gout["pred_logits"], wout["pred_logits"]),
                    "pred_embds": rel(gout["pred_embds"], wout["pred_embds"]),
                    "pred_masks_rel_l2": ((pm - wm).norm() / wm.norm()).item(),
                    "attn_mask_bits_flipped": int(flips_q.sum()), "attn_mask_bits_total": int(total_bits[0]),
                    "queries": int(nq), "queries_with_flipped_bits": int(flipped.sum()),
                    "queries_beyond_tol": int(beyond.sum()),
                    "queries_beyond_tol_without_flipped_bits": int((beyond & ~flipped).sum()),
                    "max_error_of_queries_without_flipped_bits": float(perq[~flipped].max()) if (~flipped).any() else 0.0,
                    "final_mask_sign_disagreement": ((pm < 0) != (wm < 0)).float().mean().item()},
                "same_decisions": {
                    "pred_masks": rel(fout["pred_masks"], wm), "pred_logits": rel(fout["pred_logits"], wout["pred_logits"]),
                    "pred_embds": rel(fout["pred_embds"], wout["pred_embds"])}})
            if task == "sot" and "prompt_feats" in ctg[0]:
                row["same_decisions"]["prompt_feats"] = rel(rtg[0]["prompt_feats"], ctg[0]["prompt_feats"])
                row["same_decisions"]["prompt_attn_masks_equal"] = bool(
                    torch.equal(rtg[0]["prompt_attn_masks"].cpu(), ctg[0]["prompt_attn_masks"]))
            if task == "grounding" and torch.is_tensor(fout.get("pred_reid_logits")):
                row["same_decisions"]["pred_reid_logits"] = rel(fout["pred_reid_logits"], wout["pred_reid_logits"])
            res["clips"].append(row)
            # the next clip continues from the decision-replayed state: both sides then hold the same memory pool up to
            # arithmetic error, so a flip in one clip is not counted again as a prompt difference in the next
            gtg = rtg
    finally:
        set_precision(saved_precision)
        torch.set_num_threads(old_threads)
    if not dry:
        res["pass"] = all(max(r["same_decisions"][k] for k in ("pred_masks", "pred_logits", "pred_embds")) <= TOL
                          for r in res["clips"])
    return res
