"""Tracking helpers shared by the sliding-window heads (univs/inference/comm.py:11-58, univs/utils/comm.py:85-88).

Device-first restatements: similarities are one small GEMM on the device, the only host work is the Hungarian
assignment itself (scipy, as in the reference, comm.py:53-54) on a [N, M] cost matrix copied once."""
from __future__ import annotations

import math

import torch
from scipy.optimize import linear_sum_assignment


def calculate_mask_quality_scores(mask_logits: torch.Tensor, threshold: float = 1.0) -> torch.Tensor:
    """Stability score per query (univs/utils/comm.py:85-88): #{logit > thr} / max(#{logit > -thr}, 1) over all
    trailing dims."""
    flat = mask_logits.flatten(1)
    inner = (flat > threshold).sum(-1)
    outer = (flat > -threshold).sum(-1).clamp(min=1)
    return inner / outer


def generate_temporal_weights(num_frames: int, weights: torch.Tensor | None = None, enable_softmax: bool = False,
                              scaler: float = 5.0) -> torch.Tensor:
    """Exponentially increasing weight for later frames, optionally gated per frame, normalised to sum 1
    (inference/comm.py:11-26)."""
    w = (torch.arange(1, num_frames + 1, dtype=torch.float32) / num_frames * scaler).exp()
    if enable_softmax:          # NB: softmax OF the exponentials, as the reference does (:17-19)
        w = w.softmax(-1)
    if weights is not None:
        if weights.shape[-1] != num_frames:
            raise ValueError("weights must have one entry per frame")
        w = w.to(weights) * weights
    return w / w.sum(-1, keepdim=True).clamp(min=1e-3)


def _unit(x: torch.Tensor) -> torch.Tensor:
    return x / x.norm(dim=-1, keepdim=True).clamp(min=1e-3)


def match_from_learnable_embds(tgt_embds, cur_embds, return_similarity=False, return_src_indices=False,
                               use_norm=True, thresh=0):
    """Hungarian assignment of the current clip's queries to the tracked ones (inference/comm.py:28-58).

    tgt_embds [N, V, C] (memory of V past clips), cur_embds [M, T, C].  use_norm: cosine similarity averaged over the
    clip frames, memory frames weighted by `generate_temporal_weights` (blank memory rows masked out); otherwise
    bi-softmax of scaled dot products.  Returns the permutation of current queries aligned to the targets."""
    V = tgt_embds.shape[1]
    if use_norm:
        tgt = _unit(tgt_embds)
        # mean over the clip frames commutes with the dot product: sim[n,m,v] = <tgt[n,v], mean_t cur[m,t]>
        cur = _unit(cur_embds).mean(1)
        w = generate_temporal_weights(V, weights=(tgt != 0).any(-1).float())          # [N, V]
        sim = torch.einsum("nvc,mc->nmv", tgt, cur)
        sim = (sim * w.unsqueeze(1)).sum(-1)                                          # [N, M]
    else:
        sim = torch.einsum("nvc,mc->nmv", tgt_embds, cur_embds.mean(1)) / math.sqrt(tgt_embds.shape[-1])
        sim = (sim.softmax(1) + sim.softmax(0)).mean(-1) / 2.0
        if thresh > 0:
            sim = sim.masked_fill(sim < thresh, 0.0)
    rows, cols = linear_sum_assignment((1.0 - sim).cpu().numpy())
    matched = sim[torch.as_tensor(rows, device=sim.device), torch.as_tensor(cols, device=sim.device)]
    indices = (rows, cols) if return_src_indices else cols
    return (indices, matched) if return_similarity else indices


class TemporalMaskMean:
    """Running per-frame mean of overlapping clip masks.

    The reference keeps every clip's [Q, T, h, w] logits and averages at the end
    (inference_video_vis_fast.py:273-281: frame v is the mean of out_masks[v - t][:, t] over the clips that contain
    it).  Here the sum is accumulated in place into one [Q, V, h, w] buffer (T x less memory, no second pass)."""

    def __init__(self, num_queries: int, video_len: int, size, device, dtype=torch.float32):
        self.sum = torch.zeros((num_queries, video_len, *size), device=device, dtype=dtype)
        self.count = torch.zeros(video_len, device=device, dtype=dtype)

    def add(self, start: int, clip_masks: torch.Tensor):
        """clip_masks [Q, T, h, w] for frames [start, start + T)."""
        T = clip_masks.shape[1]
        self.sum[:, start:start + T] += clip_masks
        self.count[start:start + T] += 1

    def mean(self, num_frames: int | None = None) -> torch.Tensor:
        n = self.sum.shape[1] if num_frames is None else num_frames
        return self.sum[:, :n] / self.count[:n].clamp(min=1).view(1, -1, 1, 1)


def check_consistency_with_prev_frames(prev_embds, cur_embds, sim_threshold=0.5, return_similarity=False,
                                       use_norm=True):
    """Row-wise (object i against itself) temporal consistency of prompt-query embeddings (inference/comm.py:64-95).
    prev_embds [N, V, C] memory, cur_embds [N, T, C] current clip."""
    if use_norm:
        prev = _unit(prev_embds)
        cur = _unit(cur_embds).mean(1)                                    # mean over clip frames commutes
        w = generate_temporal_weights(prev.shape[1], weights=(prev != 0).any(-1).float())
        sim = ((prev * cur.unsqueeze(1)).sum(-1) * w).sum(-1)             # [N]
        ok = sim > sim_threshold
    else:
        pair = prev_embds[:, -3:].mean(1) @ cur_embds.mean(1).t()
        pair = 0.5 * (pair.softmax(0) + pair.softmax(1))
        ok = pair.argmax(-1) == torch.arange(pair.shape[0], device=pair.device)
        sim = pair.diagonal()
        ok = ok | (sim > 0.25)
    return (ok, sim) if return_similarity else ok


def video_box_iou(boxes1, boxes2):
    """Per-frame IoU of XYXY boxes: [N, T, 4] x [M, T, 4] -> [N, M, T] (univs/utils/comm.py:137-158)."""
    area = lambda b: (b[..., 2] - b[..., 0]) * (b[..., 3] - b[..., 1])
    lt = torch.maximum(boxes1[:, None, :, :2], boxes2[None, :, :, :2])
    rb = torch.minimum(boxes1[:, None, :, 2:], boxes2[None, :, :, 2:])
    inter = (rb - lt).clamp(min=0).prod(-1)
    union = (area(boxes1)[:, None] + area(boxes2)[None] - inter).clamp(min=1e-3)
    return inter / union


def pair_mask_iou(masks1, masks2):
    """IoU of corresponding binary masks [..., H, W] -> [...] (univs/utils/comm.py:213-227)."""
    a, b = masks1.flatten(-2) > 0.5, masks2.flatten(-2) > 0.5
    return (a & b).sum(-1) / (a | b).sum(-1).clamp(min=1)


def is_semseg_dataset(dataset_name: str) -> bool:
    """univs/prepare_targets.py:13-17"""
    return dataset_name.startswith("vspw")


def process_inference(video: dict, inter_image_size, image_size, num_frames: int, semantic_on: bool = False,
                      custom_videos_text=()):
    """PrepareTargets.process_inference (univs/prepare_targets.py:46-95) for the one video of an inference batch: the
    `targets` list the decoder and the prompt sampler read and mutate.  prompt_type: "text" for grounding and for
    category-prompt detection on semantic datasets (or SEMANTIC_ON), "visual" otherwise (:58-64).  The CLIP text tower is
    not part of this build: expression features (`exp_word_feats`, `exp_sentence_feats`, `exp_word_len`) are taken from
    the input dict where the reference computes them here (:90-93, :260-330)."""
    video_in = video
    if len(custom_videos_text) > 0:
        video = dict(video, task="grounding", expressions=list(custom_videos_text[0]),
                     exp_obj_ids=list(range(len(custom_videos_text[0]))))
    task = video.get("task", "detection")
    name = video["dataset_name"]
    if task == "grounding":
        prompt_type = "text"
    elif task == "detection":
        prompt_type = "text" if (is_semseg_dataset(name) or semantic_on) else "visual"
    else:
        prompt_type = "visual"
    V = len(video["image"])
    tg = {"video_len": int(video.get("video_len", V)), "dataset_name": name, "task": task, "num_frames": num_frames,
          "inter_image_size": tuple(inter_image_size), "image_size": tuple(image_size),
          "file_names": video.get("file_names", [""] * V), "prompt_type": prompt_type}
    if "video_id" in video:
        tg["video_id"] = video["video_id"]
    for key in ("mask_palette", "expressions", "exp_obj_ids", "exp_word_feats", "exp_sentence_feats", "exp_word_len"):
        if key in video:
            tg[key] = video[key]
    if task == "sot" and "instances" in video:
        tg["instances"] = video["instances"]
    video_in["prompt_type"] = prompt_type          # the reference writes it back into the input dict as well (:81)
    return [tg]
