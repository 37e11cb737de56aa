"""Mask/box/point visual prompts -> dense prompt tokens (prompt_encoder.py:58-497, 844-1071).  Filled in below."""
from __future__ import annotations


def sample_visual_prompts(sampler, src, pos, size_list, tg):
    raise NotImplementedError("visual-prompt sampling (sot / VOS path) is not built yet in this round")
