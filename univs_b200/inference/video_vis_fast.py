"""Video instance segmentation head with the MinVIS-style frame tracker (univs/inference/inference_video_vis_fast.py).

Same contract as the reference class: `eval(model, batched_inputs)` for one video returns
{"image_size", "pred_scores", "pred_labels", "pred_masks"} (:347-352).  What differs is how the clips are produced:

* the reference re-runs the pixel decoder for every T-frame clip of the stride-1 sliding window (:223-236), i.e. T times
  per frame; here frames go through backbone + pixel decoder once (`ClipStream`) and only the decoder runs per clip;
* clip masks are averaged per frame with a running sum (`TemporalMaskMean`) instead of a list of all clip outputs
  (:257-281);
* tracking state stays on the device; the only host round trip per clip is the [100, Q] cost matrix of the Hungarian
  assignment (comm.py:53-54 does the same).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from ..modeling.decoder import COMBINED_DATASETS_CATEGORY_INFO
from ..registry import is_cfg
from ..streaming import ClipStream
from . import rle
from .comm import TemporalMaskMean, calculate_mask_quality_scores, match_from_learnable_embds, process_inference


class InferenceVideoVISFast(nn.Module):
    def __init__(self, cfg=None, *, num_queries=200, num_frames=5, size_divisibility=32, stability_score_thresh=0.0,
                 test_topk_per_image=100, zero_shot_inference=False, tracker_type="minvis", merge_on_cpu=False,
                 num_frames_window_test=5, lsj_aug_enable_test=False, lsj_aug_image_size=1024, reuse_features=True,
                 rle_output=False):
        super().__init__()
        if cfg is not None and is_cfg(cfg):
            mf, bv = cfg.MODEL.MASK_FORMER, cfg.MODEL.BoxVIS.TEST
            num_queries = mf.NUM_OBJECT_QUERIES
            num_frames = cfg.INPUT.SAMPLING_FRAME_NUM
            size_divisibility = mf.SIZE_DIVISIBILITY
            stability_score_thresh = mf.TEST.get("STABILITY_SCORE_THRESH", 0.0)
            test_topk_per_image = cfg.get("TEST", {}).get("DETECTIONS_PER_IMAGE", 100)
            zero_shot_inference = bv.get("ZERO_SHOT_INFERENCE", False)
            tracker_type = bv.get("TRACKER_TYPE", "minvis")
            merge_on_cpu = bv.get("MERGE_ON_CPU", False)
            num_frames_window_test = bv.NUM_FRAMES_WINDOW
            lsj_aug_enable_test = cfg.INPUT.LSJ_AUG.SQUARE_ENABLED
            lsj_aug_image_size = cfg.INPUT.LSJ_AUG.IMAGE_SIZE
        self.num_queries = num_queries
        self.num_frames = num_frames
        self.size_divisibility = size_divisibility
        self.stability_score_thresh = stability_score_thresh
        self.test_topk_per_image = test_topk_per_image
        self.zero_shot_inference = zero_shot_inference
        self.tracker_type = tracker_type
        self.merge_on_cpu = merge_on_cpu
        self.num_frames_window_test = max(num_frames_window_test, num_frames)
        self.LSJ_aug_enable_test = lsj_aug_enable_test
        self.LSJ_aug_image_size = lsj_aug_image_size
        self.reuse_features = reuse_features
        # True: return COCO RLE per (instance, frame) -- what the reference's writers build on the host from the dense masks
        # (inference_video_vis.py:526-531) -- with the run scan on the device, instead of the dense masks themselves
        self.rle_output = rle_output

    # ------------------------------------------------------------------ entry point (reference :185-218)
    @torch.no_grad()
    def eval(self, model, batched_inputs):
        if len(batched_inputs) != 1:
            raise ValueError("one video per call")
        video = batched_inputs[0]
        dataset_name = video["dataset_name"]
        if not (dataset_name.startswith("ytvis") or dataset_name.startswith("ovis")):
            raise ValueError(f"Do not support the model inference on {dataset_name}.")
        if self.tracker_type != "minvis":
            raise ValueError("the type of tracker only supports minvis.")
        x, image_size = model.preprocess(video["image"])
        if self.LSJ_aug_enable_test:      # ImageList.from_tensors(..., square_size): pad to a fixed square
            d, S = self.size_divisibility, self.LSJ_aug_image_size
            S = (max(S, *x.shape[-2:]) + d - 1) // d * d
            x = F.pad(x, (0, S - x.shape[-1], 0, S - x.shape[-2]), value=0.0)
        targets = video.get("targets")
        if targets is None:
            # detection on a VIS dataset: visual prompt type without masks => learnable queries only
            # (prepare_targets.py:58-64, prompt_encoder.py:809-810)
            targets = process_inference(video, tuple(x.shape[-2:]), image_size, self.num_frames)
        images = _Images(x, [image_size] * x.shape[0])
        return self.inference_video_vis_minvis(model, batched_inputs, images, targets)

    # ------------------------------------------------------------------ clip loop + tracker (reference :220-296)
    @torch.no_grad()
    def inference_video_vis_minvis(self, model, batched_inputs, images, targets):
        x = images.tensor
        V, T, Q = x.shape[0], self.num_frames, self.num_queries
        if V < T:
            raise ValueError(f"video of {V} frames is shorter than one clip ({T})")
        dataset_name = batched_inputs[0]["dataset_name"]
        if dataset_name not in COMBINED_DATASETS_CATEGORY_INFO:
            raise KeyError(dataset_name)
        num_classes, first_class = COMBINED_DATASETS_CATEGORY_INFO[dataset_name]

        stream = ClipStream(model, T, max_cached_frames=2 * max(T, self.num_frames_window_test)) \
            if self.reuse_features else None
        pushed, window = 0, (0, 0, None)
        logit_sum, masks, memory, n_clips = None, None, [], 0
        for i in range(V - T + 1):
            if stream is not None:
                while pushed < i + T:       # encode each frame once, a window-sized batch at a time
                    k = min(self.num_frames_window_test, V - pushed)
                    stream.push_preprocessed(pushed, x[pushed:pushed + k])
                    pushed += k
                out = stream.clip(i, targets)
            else:                           # reference schedule: backbone per window, pixel decoder per clip
                if i + T > window[1]:
                    window = (i, i + self.num_frames_window_test, model.backbone(x[i:i + self.num_frames_window_test]))
                feats = {k: v[i - window[0]:i - window[0] + T] for k, v in window[2].items()}
                targets[0]["frame_indices"] = torch.arange(i, i + T)
                out = model.sem_seg_head(feats, targets=targets)
            out = {k: v for k, v in out.items() if torch.is_tensor(v)}       # drops aux_outputs (:237)

            if i == 0:
                # keep the 100 most confident queries of the first clip as the tracks (:239-245)
                scores = out["pred_logits"][0].sigmoid()
                if self.stability_score_thresh > 0:
                    scores = scores + calculate_mask_quality_scores(out["pred_masks"][0]).view(-1, 1)
                keep = scores.max(-1)[0].sort(descending=True)[1][:min(Q, 100)]
                out = {k: v[:, keep] for k, v in out.items()}
            if self.merge_on_cpu:
                out = {k: v.cpu() for k, v in out.items()}
            logits = out["pred_logits"][0, :Q].float()       # [q, K]
            clip_masks = out["pred_masks"][0, :Q].float()     # [q, T, h, w]
            embds = out["pred_embds"][0, :Q].float()          # [q, T, C]
            if i > 0:
                # align this clip's queries with the tracks: memory = mean embeddings of the last two clips (:256-258)
                order = match_from_learnable_embds(torch.stack(memory[-2:], dim=1), embds)
                order = torch.as_tensor(order, device=embds.device)
                logits, clip_masks, embds = logits[order], clip_masks[order], embds[order]
            else:
                logit_sum = torch.zeros_like(logits)
                masks = TemporalMaskMean(clip_masks.shape[0], V, clip_masks.shape[-2:], clip_masks.device)
            logit_sum += logits
            masks.add(i, clip_masks)
            memory = memory[-1:] + [embds.mean(1)]
            n_clips += 1

        outputs = {
            "pred_masks": masks.mean(),                                                        # [q, V, h, w]
            "pred_scores": (logit_sum / n_clips)[..., first_class:first_class + num_classes].sigmoid(),   # [q, k]
        }
        interim_size = tuple(x.shape[-2:])
        image_size = images.image_sizes[0]
        out_size = (batched_inputs[0].get("height", image_size[0]), batched_inputs[0].get("width", image_size[1]))
        return self.inference_video_vis_minvis_save_video(model, images, outputs, interim_size, image_size, out_size)

    # ------------------------------------------------------------------ scoring + resize to the output size (:298-354)
    @torch.no_grad()
    def inference_video_vis_minvis_save_video(self, model, images, outputs, interim_size, image_size, out_size):
        scores, mask_pred = outputs["pred_scores"], outputs["pred_masks"]
        top = scores.max(-1)[0].sort(descending=True)[1][:self.test_topk_per_image]
        scores, mask_pred = scores[top], mask_pred[top]
        if self.zero_shot_inference:
            scores = (scores * 20).softmax(-1)
        K = scores.shape[-1]
        # (query, class) pairs scoring above twice the uniform level, at least 5, at most topk
        n_keep = min(self.test_topk_per_image, max(int((scores > 2.0 / K).sum()), 5))
        scores_per_video, flat = scores.flatten().topk(n_keep, sorted=False)
        labels_per_video = flat % K
        query_of = torch.div(flat, K, rounding_mode="floor")

        mask_pred = F.interpolate(mask_pred[query_of], size=interim_size, mode="bilinear", align_corners=False)
        mask_pred = mask_pred[:, :, :image_size[0], :image_size[1]]
        step = max(int(mask_pred.shape[1] / 10.0), 1)
        quality = calculate_mask_quality_scores(mask_pred[:, ::step]).clamp(min=0.1)
        scores_per_video = scores_per_video * quality.to(scores_per_video.device)

        masks_per_video, rles_per_video = [], []
        for m in mask_pred:        # one object at a time: bounded memory for long videos (:330-338)
            m = F.interpolate(m.unsqueeze(0), size=out_size, mode="bilinear", align_corners=False).squeeze(0) > 0.0
            if self.rle_output:
                rles_per_video.append(rle.encode(m))           # [V] dicts; only the run boundaries leave the device
            else:
                masks_per_video.append(m.cpu())
        out = {"image_size": out_size, "pred_scores": scores_per_video.tolist(),
               "pred_labels": labels_per_video.tolist(), "pred_masks": masks_per_video}
        if self.rle_output:
            out["segmentations"] = rles_per_video
        return out


class _Images:
    """The two ImageList fields the heads read (detectron2.structures.ImageList: .tensor, .image_sizes)."""

    def __init__(self, tensor, image_sizes):
        self.tensor = tensor
        self.image_sizes = image_sizes
