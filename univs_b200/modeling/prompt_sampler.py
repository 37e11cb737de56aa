"""Visual prompt sampler (inference), host side.

Mirror of univs/modeling/prompt_encoder/prompt_encoder.py::VisualPromptSampler (:499-1071) for the inference entry
`process_per_batch` (:529-541 -> process_per_batch_inference :781-842).  The per-clip prompt memory pool lives in
`targets[0]` (`prompt_feats`, `prompt_pe`, `prompt_attn_masks`) exactly as in the reference; ownership stays with
the caller.
"""
from __future__ import annotations

import torch


class VisualPromptSampler:
    def __init__(self, pretrain_img_size=1024, hidden_dim=256, num_heads=8, num_frames=1, num_prev_frames_memory=1,
                 num_dense_points=32, position_embedding_sin3d_type="FixedT", clip_stride=1):
        self.num_heads = num_heads
        self.num_frames = num_frames
        self.key_fid = int((num_frames - 1) / 2)
        self.num_dense_points = num_dense_points
        self.clip_stride = clip_stride
        self.num_prev_frames_memory = max(num_prev_frames_memory, num_frames)
        self.hidden_dim = hidden_dim
        self.pretrain_img_size = pretrain_img_size
        self.position_embedding_sin3d_type = position_embedding_sin3d_type
        self.prompt_feature_level_index = -1        # the 1/8-scale level (prompt_encoder.py:526)
        self.img_feats_scale = 8                    # prompt_encoder.py:79

    @torch.no_grad()
    def process_per_batch(self, src, pos, size_list, targets):
        """src/pos: lists of [T,S_l,C] token-major level memories.  Returns (prompt_pe_dense, prompt_feats_dense),
        each [P, R, T, C], or (None, None) when the clip carries no visual prompts (prompt_encoder.py:809-810)."""
        assert len(targets) == 1, "Only support batch size = 1 now"
        tg = targets[0]
        if "masks" not in tg or tg["masks"].nelement() == 0:
            return None, None
        from .visual_prompts import sample_visual_prompts
        return sample_visual_prompts(self, src, pos, size_list, tg)

    @torch.no_grad()
    def memory_pool_prompts(self, tg, num_prev_frames_memory):
        """extract_prompt_features_from_memoey_pool (..._univs.py:795-822) without the T-fold repeat:
        returns (pe, feats) as [P, 1, R*(1+T_prev'), C] (T-invariant memory, frame stride 0 in ProCA)."""
        pf, pp = tg["prompt_feats"], tg["prompt_pe"]                 # [P,R,n_frames,C]
        P, _, e_idx = pf.shape[:3]
        first = tg["first_appear_frame_idxs"].clone()
        first[first >= e_idx - 1] = -1
        ar = torch.arange(P, device=pf.device)
        f_first, p_first = pf[ar, :, first], pp[ar, :, first]       # [P,R,C]
        f_prev = pf[:, :, -num_prev_frames_memory:].transpose(1, 2).flatten(1, 2)
        p_prev = pp[:, :, -num_prev_frames_memory:].transpose(1, 2).flatten(1, 2)
        feats = torch.cat([f_first, f_prev], 1).unsqueeze(1).contiguous()
        pe = torch.cat([p_first, p_prev], 1).unsqueeze(1).contiguous()
        return pe, feats
