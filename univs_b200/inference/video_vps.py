"""Video panoptic segmentation head, online with the MinVIS-style frame tracker (univs/inference/inference_video_vps.py).

Same contract as the reference class: `eval(model, batched_inputs)` for one VIPSeg video returns
{"image_size", "pred_masks" (int32 [V, H, W] segment ids, on the host), "segments_infos", "pred_ids", "task": "vps"}
(:399-406).  As in the VIS head, frames go through backbone + pixel decoder once (`ClipStream`) instead of once per
clip (:223-236), and the per-frame mask mean is a running sum instead of a list of all clip outputs (:262-270); the
tracker (:295-307) and the panoptic assembly (:309-406) follow the reference step by step.  `thing_ids` replaces
`metadata.thing_dataset_id_to_contiguous_id.keys()` (the 1-based dataset ids of the thing classes)."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from ..modeling.decoder import COMBINED_DATASETS_CATEGORY_INFO
from ..registry import is_cfg
from ..streaming import ClipStream
from .comm import TemporalMaskMean, calculate_mask_quality_scores, process_inference


def match_from_embds(tgt_embds, cur_embds):
    """inference_video_vps.py:295-307: Hungarian assignment on 1 - cosine similarity; returns, for every target row, the
    index of the current row aligned to it."""
    from scipy.optimize import linear_sum_assignment
    cur = cur_embds / cur_embds.norm(dim=1)[:, None]
    tgt = tgt_embds / tgt_embds.norm(dim=1)[:, None]
    cost = (1 - cur @ tgt.t()).cpu()
    return linear_sum_assignment(cost.transpose(0, 1))[1]


class InferenceVideoVPS(nn.Module):
    def __init__(self, cfg=None, *, num_queries=200, num_frames=5, size_divisibility=32, object_mask_threshold=0.05,
                 overlap_threshold=0.8, stability_score_thresh=0.0, test_topk_per_image=100, merge_on_cpu=False,
                 num_frames_window_test=5, lsj_aug_enable_test=False, lsj_aug_image_size=1024, thing_ids=(),
                 change_to_720p=True, reuse_features=True):
        super().__init__()
        if cfg is not None and is_cfg(cfg):
            mf, bv = cfg.MODEL.MASK_FORMER, cfg.MODEL.BoxVIS.TEST
            num_queries = mf.NUM_OBJECT_QUERIES
            num_frames = cfg.INPUT.SAMPLING_FRAME_NUM
            size_divisibility = mf.SIZE_DIVISIBILITY
            object_mask_threshold = mf.TEST.OBJECT_MASK_THRESHOLD
            overlap_threshold = mf.TEST.OVERLAP_THRESHOLD
            stability_score_thresh = mf.TEST.get("STABILITY_SCORE_THRESH", 0.0)
            test_topk_per_image = cfg.get("TEST", {}).get("DETECTIONS_PER_IMAGE", 100)
            merge_on_cpu = bv.get("MERGE_ON_CPU", False)
            num_frames_window_test = bv.NUM_FRAMES_WINDOW
            lsj_aug_enable_test = cfg.INPUT.LSJ_AUG.SQUARE_ENABLED
            lsj_aug_image_size = cfg.INPUT.LSJ_AUG.IMAGE_SIZE
        self.num_queries, self.num_frames = num_queries, num_frames
        self.size_divisibility = size_divisibility
        self.object_mask_threshold = object_mask_threshold
        self.overlap_threshold = overlap_threshold
        self.stability_score_thresh = stability_score_thresh
        self.test_topk_per_image = test_topk_per_image
        self.merge_on_cpu = merge_on_cpu
        self.num_frames_window_test = max(num_frames_window_test, num_frames)
        self.LSJ_aug_enable_test, self.LSJ_aug_image_size = lsj_aug_enable_test, lsj_aug_image_size
        self.thing_ids = set(int(i) for i in thing_ids)
        self.change_to_720p = change_to_720p
        self.reuse_features = reuse_features

    # ------------------------------------------------------------------ entry point (reference :175-207)
    @torch.no_grad()
    def eval(self, model, batched_inputs):
        if len(batched_inputs) != 1:
            raise ValueError("one video per call")
        video = batched_inputs[0]
        dataset_name = video["dataset_name"]
        if not dataset_name.startswith("vipseg"):
            raise ValueError(f"Not support to eval {dataset_name} during training yet.")
        x, image_size = model.preprocess(video["image"])
        if self.LSJ_aug_enable_test:
            d, S = self.size_divisibility, self.LSJ_aug_image_size
            S = (max(S, *x.shape[-2:]) + d - 1) // d * d
            x = F.pad(x, (0, S - x.shape[-1], 0, S - x.shape[-2]), value=0.0)
        targets = video.get("targets")
        if targets is None:
            targets = process_inference(video, tuple(x.shape[-2:]), image_size, self.num_frames)
        return self.inference_video_vps_online(model, batched_inputs, x, image_size, targets)

    # ------------------------------------------------------------------ clip loop + tracker (reference :209-293)
    @torch.no_grad()
    def inference_video_vps_online(self, model, batched_inputs, x, image_size, targets):
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
                while pushed < i + T:
                    k = min(self.num_frames_window_test, V - pushed)
                    stream.push_preprocessed(pushed, x[pushed:pushed + k])
                    pushed += k
                out = stream.clip(i, targets)
            else:
                if i + T > window[1]:
                    window = (i, i + self.num_frames_window_test, model.backbone(x[i:i + self.num_frames_window_test]))
                feats = {k: v[i - window[0]:i - window[0] + T] for k, v in window[2].items()}
                targets[0]["frame_indices"] = torch.arange(i, i + T)
                out = model.sem_seg_head(feats, targets=targets)
            out = {k: v for k, v in out.items() if torch.is_tensor(v)}
            if i == 0:
                scores = out["pred_logits"][0].sigmoid()
                if self.stability_score_thresh > 0:
                    scores = scores + calculate_mask_quality_scores(out["pred_masks"][0]).view(-1, 1)
                keep = scores.max(-1)[0].sort(descending=True)[1][:min(Q, 100)]
                out = {k: v[:, keep] for k, v in out.items()}
            if self.merge_on_cpu:
                out = {k: v.cpu() for k, v in out.items()}
            logits = out["pred_logits"][0, :Q].float()                 # [q, K]
            clip_masks = out["pred_masks"][0, :Q].float()               # [q, T, h, w]
            embds = out["pred_embds"][0, :Q].float().mean(1)            # [q, C]
            if i > 0:
                order = torch.as_tensor(match_from_embds(torch.stack(memory[-2:]).mean(0), embds), device=embds.device)
                logits, clip_masks, embds = logits[order], clip_masks[order], embds[order]
            else:
                logit_sum = torch.zeros_like(logits)
                masks = TemporalMaskMean(clip_masks.shape[0], V, clip_masks.shape[-2:], clip_masks.device)
            logit_sum += logits
            masks.add(i, clip_masks)
            memory = memory[-1:] + [embds]
            n_clips += 1
        pred_cls = (logit_sum / n_clips)[..., first_class:first_class + num_classes].sigmoid()
        interim_size = tuple(x.shape[-2:])
        out_h, out_w = batched_inputs[0].get("height", image_size[0]), batched_inputs[0].get("width", image_size[1])
        out_size = (720, int(720 * out_w / out_h)) if self.change_to_720p else (out_h, out_w)
        return self.inference_video_vps_save_results(pred_cls, masks.mean(), interim_size, image_size, out_size)

    # ------------------------------------------------------------------ panoptic assembly (reference :309-406)
    @torch.no_grad()
    def inference_video_vps_save_results(self, pred_cls, pred_masks, interim_size, img_size, out_size):
        scores, labels = pred_cls.max(-1)
        pred_id = torch.arange(len(scores), device=pred_cls.device)
        keep = scores > max(self.object_mask_threshold, float(scores.topk(k=self.test_topk_per_image)[0][-1]))
        cur_scores, cur_classes, cur_masks, cur_ids = scores[keep], labels[keep], pred_masks[keep], pred_id[keep]
        panoptic_seg = torch.zeros((cur_masks.size(1), out_size[0], out_size[1]), dtype=torch.int32, device=cur_masks.device)
        segments_infos, out_ids = [], []
        if cur_masks.shape[0] > 0:
            t_itv = 10
            cur_masks = torch.cat([
                F.interpolate(cur_masks[:, t:t + t_itv], size=interim_size, mode="bilinear", align_corners=False)[
                    :, :, :img_size[0], :img_size[1]] for t in range(0, cur_masks.shape[1], t_itv)], dim=1)
            cur_scores = cur_scores + 0.5 * calculate_mask_quality_scores(cur_masks[:, ::5])
            cur_masks = cur_masks.sigmoid()
            is_bg = (cur_masks < 0.5).sum(0) == len(cur_masks)
            cur_mask_ids = (cur_scores.view(-1, 1, 1, 1).to(cur_masks.device) * cur_masks).argmax(0)     # [t, h, w]
            cur_mask_ids[is_bg] = -1
            cur_mask_ids = F.interpolate(cur_mask_ids.float().unsqueeze(0), size=out_size, mode="nearest").long().squeeze(0)
            stuff_memory, current_segment_id = {}, 0
            for k in range(cur_classes.shape[0]):
                m_k = F.interpolate(cur_masks[k].unsqueeze(0), size=out_size, mode="bilinear", align_corners=False).squeeze(0)
                pred_class = int(cur_classes[k]) + 1                  # dataset ids start from 1
                isthing = pred_class in self.thing_ids
                own = cur_mask_ids == k
                mask_area = int(own.sum())
                original_area = int((m_k >= 0.5).sum())
                mask = own & (m_k >= 0.5)
                if mask_area > 0 and original_area > 0 and int(mask.sum()) > 0:
                    if mask_area / original_area < self.overlap_threshold:
                        continue
                    if not isthing:                                  # merge stuff regions of one class
                        if pred_class in stuff_memory:
                            panoptic_seg[mask] = stuff_memory[pred_class]
                            continue
                        stuff_memory[pred_class] = current_segment_id + 1
                    current_segment_id += 1
                    panoptic_seg[mask] = current_segment_id
                    segments_infos.append({"id": current_segment_id, "isthing": bool(isthing), "category_id": pred_class})
                    out_ids.append(cur_ids[k])
        return {"image_size": out_size, "pred_masks": panoptic_seg.cpu(), "segments_infos": segments_infos,
                "pred_ids": out_ids, "task": "vps"}
