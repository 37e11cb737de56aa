"""Sliding-window task heads over the hot path (callers of `model.backbone` / `model.sem_seg_head`,
univs/inference/*).  Built on `ClipStream`: every frame is encoded once."""
from .comm import (calculate_mask_quality_scores, generate_temporal_weights, match_from_learnable_embds,
                   TemporalMaskMean)
from .video_vis_fast import InferenceVideoVISFast

__all__ = ["InferenceVideoVISFast", "match_from_learnable_embds", "generate_temporal_weights",
           "calculate_mask_quality_scores", "TemporalMaskMean"]
