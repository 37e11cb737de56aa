"""Builds the UniVS_Prompt hot-path model from a cfg (the reference's `Trainer.build_model(cfg)` role)."""
from __future__ import annotations

import torch

from . import modeling  # noqa: F401  (registers the classes)
from .config import SWIN_VARIANTS, get_cfg
from .meta_arch import UniVS_Prompt


def make_cfg(variant="large", num_queries=200, num_frames=5, clip_emb=None, **univs_overrides):
    cfg = get_cfg()
    for k, v in SWIN_VARIANTS[variant].items():
        cfg.MODEL.SWIN[k] = v
    cfg.MODEL.MASK_FORMER.NUM_OBJECT_QUERIES = num_queries
    cfg.INPUT.SAMPLING_FRAME_NUM = num_frames
    if clip_emb is not None:
        cfg.MODEL.UniVS.CLIP_CLASS_EMBED_PATH = clip_emb          # a tensor is accepted in place of the .pth path
    for k, v in univs_overrides.items():
        cfg.MODEL.UniVS[k] = v
    return cfg


def build_model(cfg, process_group=None, seed=0, perturb_msda=True):
    """Random-init model (there are no checkpoints offline).  MSDeformAttn's sampling_offsets / attention_weights
    linears are zero-initialised by the reference (ms_deform_attn.py:67,75), which makes sampling input-independent;
    `perturb_msda` gives them N(0, 0.02^2) weights so the gather path is data-dependent."""
    torch.manual_seed(seed)
    model = UniVS_Prompt(cfg, process_group=process_group)
    if perturb_msda:
        g = torch.Generator().manual_seed(seed + 1)
        for layer in model.sem_seg_head.pixel_decoder.transformer.encoder.layers:
            a = layer.self_attn
            with torch.no_grad():
                a.sampling_offsets.weight.copy_(torch.randn(a.sampling_offsets.weight.shape, generator=g) * 0.02)
                a.attention_weights.weight.copy_(torch.randn(a.attention_weights.weight.shape, generator=g) * 0.02)
    return model
