"""Semantic-extraction head (univs/inference/inference_video_semantic_extraction.py): object tokens and spatially
compressed mask features of a whole video, for downstream video-language models.

Same contract as the reference class: `eval(model, batched_inputs)` for one video walks non-overlapping clips of T frames
(the decoder must be built with `semantic_extraction_enable=True`, so that it returns the un-normalised query tokens
[T, C, Q] and the mask features [T, C, h, w], ..._univs.py:448-452), resizes the mask features to the padded input size,
crops the padding, samples them down (nearest) by COMPRESSION_RATIO relative to the output size, and writes
`<video_id>._obj_tokens_<s>_<t>.pt` / `<video_id>._compression_mask_features_<s>_<t>.pt` (:249-261).  Here the two tensors
are also returned ({"obj_tokens": [V', C, Q], "compression_mask_features": [V', C, H/s, W/s]}); files are written only
when an output directory is configured or derivable from the file names ("raw" -> "semantic_extraction", :251-255).
Clips do not overlap, so there is nothing to reuse across clips: frames go through the hot path once either way."""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F
from torch import nn

from ..registry import is_cfg
from .comm import process_inference


class InferenceVideoSemanticExtraction(nn.Module):
    def __init__(self, cfg=None, *, num_frames=5, size_divisibility=32, compression_ratio=32, compression_ratio_temporal=1,
                 output_dir="", lsj_aug_enable_test=False, lsj_aug_image_size=1024, save=True):
        super().__init__()
        if cfg is not None and is_cfg(cfg):
            se = cfg.MODEL.UniVS.TEST.SEMANTIC_EXTRACTION
            num_frames = cfg.INPUT.SAMPLING_FRAME_NUM
            size_divisibility = cfg.MODEL.MASK_FORMER.SIZE_DIVISIBILITY
            compression_ratio = se.get("COMPRESSION_RATIO", 32)
            compression_ratio_temporal = se.get("COMPRESSION_RATIO_TEMPORAL", 1)
            output_dir = se.get("OUTPUT_DIR", "")
            lsj_aug_enable_test = cfg.INPUT.LSJ_AUG.SQUARE_ENABLED
            lsj_aug_image_size = cfg.INPUT.LSJ_AUG.IMAGE_SIZE
        self.num_frames = num_frames
        self.num_frames_window_test = 2 * num_frames                      # :85
        self.size_divisibility = size_divisibility
        self.compression_ratio = compression_ratio
        self.compression_ratio_temporal = compression_ratio_temporal
        self.output_dir = output_dir
        self.LSJ_aug_enable_test, self.LSJ_aug_image_size = lsj_aug_enable_test, lsj_aug_image_size
        self.save = save

    @torch.no_grad()
    def eval(self, model, batched_inputs):
        if len(batched_inputs) != 1:
            raise ValueError("one video per call")
        video = batched_inputs[0]
        if not getattr(model.sem_seg_head.predictor, "semantic_extraction_enable", False):
            raise RuntimeError("semantic extraction needs a decoder built with MODEL.UniVS.TEST.SEMANTIC_EXTRACTION.ENABLE")
        x, image_size = model.preprocess(video["image"])
        if self.LSJ_aug_enable_test:
            d, S = self.size_divisibility, self.LSJ_aug_image_size
            S = (max(S, *x.shape[-2:]) + d - 1) // d * d
            x = F.pad(x, (0, S - x.shape[-1], 0, S - x.shape[-2]), value=0.0)
        V = x.shape[0]
        targets = video.get("targets")
        if targets is None:
            targets = process_inference(video, tuple(x.shape[-2:]), image_size, self.num_frames)
        return self.inference_video(model, batched_inputs, x, image_size, targets)

    @torch.no_grad()
    def inference_video(self, model, batched_inputs, x, image_size, targets):
        video = batched_inputs[0]
        V, T = x.shape[0], self.num_frames
        video_id = video["video_id"]
        interim_size = tuple(x.shape[-2:])
        out_h, out_w = video.get("height", image_size[0]), video.get("width", image_size[1])
        small = (int(out_h / self.compression_ratio), int(out_w / self.compression_ratio))
        tokens, feats = [], []
        window, is_last = (0, 0, None), False
        for i in range(0, V, T):                                           # stride = clip length (:198)
            if is_last and i + T > V:
                break
            is_last = i + T >= V
            targets[0]["first_frame_idx"] = i
            targets[0]["frame_indices"] = torch.arange(i, min(i + T, V))
            if i + T > window[1]:
                window = (i, i + self.num_frames_window_test, model.backbone(x[i:i + self.num_frames_window_test]))
            clip = {k: v[i - window[0]:i - window[0] + T] for k, v in window[2].items()}
            out = model.sem_seg_head(clip, targets=targets)
            mf = F.interpolate(out["mask_features"], size=interim_size, mode="bilinear", align_corners=False)
            mf = mf[..., :image_size[0], :image_size[1]]
            tokens.append(out["pred_embds"])                                # [t, C, Q]
            feats.append(F.interpolate(mf, size=small, mode="nearest"))     # [t, C, H/s, W/s]
        tokens, feats = torch.cat(tokens), torch.cat(feats)
        assert int(video.get("video_len", V)) == tokens.shape[0]
        s_itv, t_itv = self.compression_ratio, self.compression_ratio_temporal
        result = {"obj_tokens": tokens[::t_itv], "compression_mask_features": feats[::t_itv], "task": "semantic_extraction"}
        out_dir = self.output_dir
        if not out_dir:
            names = targets[0].get("file_names") or [""]
            video_path = "/".join(names[0].split("/")[:-2])
            out_dir = video_path.replace("raw", "semantic_extraction") if "raw" in video_path else ""
        if self.save and out_dir:
            os.makedirs(out_dir, exist_ok=True)
            torch.save(result["obj_tokens"], os.path.join(out_dir, f"{video_id}._obj_tokens_{s_itv}_{t_itv}.pt"))
            torch.save(result["compression_mask_features"],
                       os.path.join(out_dir, f"{video_id}._compression_mask_features_{s_itv}_{t_itv}.pt"))
            result["files"] = out_dir
        return result
