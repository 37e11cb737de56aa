"""UniVS prompt-as-query masked transformer decoder, B200-native host side (inference).

Drop-in for univs/modeling/transformer_decoder/video_mask2former_transformer_decoder_univs.py::
VideoMultiScaleMaskedTransformerDecoderUniVS (:27-848): same ctor kwargs / from_config, same state_dict keys
(SURVEY.md App. B), `forward(x, mask_features, mask_features_bfe_conv, mask, targets)` returns the same dict
(`pred_logits`, `pred_masks [B,Q,T,H/4,W/4]`, `pred_embds`, `pred_reid_logits`, `aux_outputs`).

What is different underneath (per clip, batch = 1 video):
  * tokens are kept frame-major [T, Q, C]; memory levels are [T, S_l, C] token-major (no permutes per layer);
  * the K/V in-projections of each cross-attention layer read the level memory once; the per-head boolean mask
    [T*8, Q, S] and its float temporaries (:555-566) are replaced by a bit mask [T, Q, S/32] + a row flag produced
    directly from the mask logits by `attn_mask_bits`; the "fully blocked row" fix (:390) is that flag;
  * masked cross-attention, Q*T self-attention and ProCA run in the hand-written kernels of csrc/mha.cu;
  * the mask einsum reads channel-last mask features; `pred_masks` is its output buffer viewed as [1,Q,T,H,W];
  * `aux_outputs` (deleted by every inference caller, e.g. inference_video_vis_fast.py:237) is only materialised
    when `return_aux_outputs=True`.
Training-only branches (randperm of mask_embed :524, reid logits :531-533, stage-3 prompt merging :742-749) are
out of scope and raise.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import nn_ops, ops, switches
from ..registry import TRANSFORMER_DECODER_REGISTRY, is_cfg
from . import position
from .prompt_sampler import VisualPromptSampler

# datasets/concept_emb/combined_datasets_category_info.py:7-24 -- dataset -> (num_classes, start row) into the
# 3938-row CLIP class-embedding table
COMBINED_DATASETS_CATEGORY_INFO = {
    "imagenet": (1000, 0), "lvis": (1203, 1000), "burst": (1203, 1000), "ytvis21": (40, 2203), "ovis": (25, 2243),
    "bdd_track": (8, 2268), "objects365": (365, 2276), "coco_panoptic": (133, 2641), "coco": (80, 2641),
    "ade20k": (150, 2774), "vipseg": (124, 2924), "vspw": (124, 2924), "viposeg": (124, 2924),
    "ytvis19": (40, 3048), "entityseg_instance": (206, 3088), "entityseg_panoptic": (644, 3294),
}


class _PackedMHA(nn.Module):
    """Parameter layout of nn.MultiheadAttention (in_proj_weight [3C,C], in_proj_bias, out_proj.*)."""

    def __init__(self, d_model):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d_model, d_model))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = nn.Linear(d_model, d_model)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.out_proj.weight)
        nn.init.zeros_(self.out_proj.bias)
        self.d = d_model

    def wq(self):
        return self.in_proj_weight[: self.d], self.in_proj_bias[: self.d]

    def wk(self):
        return self.in_proj_weight[self.d: 2 * self.d], self.in_proj_bias[self.d: 2 * self.d]

    def wv(self):
        return self.in_proj_weight[2 * self.d:], self.in_proj_bias[2 * self.d:]

    def wqk(self):
        return self.in_proj_weight[: 2 * self.d], self.in_proj_bias[: 2 * self.d]

    def folded_out_bias(self):
        """out_proj.bias + out_proj.weight @ b_v : softmax rows sum to one, so the value bias passes through the
        attention unchanged and can be added after the output projection (cached per parameter version)."""
        key = (self.in_proj_bias._version, self.out_proj.weight._version, self.out_proj.bias._version,
               self.in_proj_bias.data_ptr())
        if getattr(self, "_fold", None) is None or self._fold[0] != key:
            with torch.no_grad():
                bv = self.in_proj_bias[2 * self.d:].double()
                fb = (self.out_proj.bias.double() + self.out_proj.weight.double() @ bv).float().contiguous()
            self._fold = (key, fb)
        return self._fold[1]


class _SelfAttnLayer(nn.Module):          # transformer_layers.py:11-66 (post-norm)
    def __init__(self, d):
        super().__init__()
        self.self_attn = _PackedMHA(d)
        self.norm = nn.LayerNorm(d)


class _CrossAttnLayer(nn.Module):         # transformer_layers.py:69-148 (post-norm)
    def __init__(self, d):
        super().__init__()
        self.multihead_attn = _PackedMHA(d)
        self.norm = nn.LayerNorm(d)


class _FFNLayer(nn.Module):               # transformer_layers.py:151-191 (post-norm)
    def __init__(self, d, dff):
        super().__init__()
        self.linear1 = nn.Linear(d, dff)
        self.linear2 = nn.Linear(dff, d)
        self.norm = nn.LayerNorm(d)
        nn.init.xavier_uniform_(self.linear1.weight)
        nn.init.xavier_uniform_(self.linear2.weight)

    def forward(self, x):
        z = nn_ops.linear_prepped(nn_ops.linear_act_operand(nn_ops.prep(x), self.linear1), self.linear2.weight, None)
        return nn_ops.layernorm(x, self.norm, residual=z, for_gemm=False, residual_bias=self.linear2.bias)[1]


class _MLP(nn.Module):                    # transformer_layers.py:205-217
    def __init__(self, din, dh, dout, n):
        super().__init__()
        dims = [din] + [dh] * (n - 1) + [dout]
        self.layers = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))

    def forward(self, x, prepped=False):
        h = x if prepped else nn_ops.prep(x)
        for i, l in enumerate(self.layers):
            if i < len(self.layers) - 1:
                h = nn_ops.linear_act_operand(h, l)
            else:
                y = nn_ops.linear_prepped(h, l.weight, l.bias)
        return y


@TRANSFORMER_DECODER_REGISTRY.register()
class VideoMultiScaleMaskedTransformerDecoderUniVS(nn.Module):
    _version = 2

    def __init__(self, in_channels, mask_classification=True, *args, **kwargs):
        super().__init__()
        if is_cfg(in_channels):               # @configurable: (cfg, in_channels, mask_classification)
            cfg = in_channels
            kwargs = self.from_config(cfg, mask_classification, *args)
            in_channels = kwargs.pop("in_channels")
            mask_classification = kwargs.pop("mask_classification")
        self._build(in_channels, mask_classification, **kwargs)

    @classmethod
    def from_config(cls, cfg, in_channels, mask_classification):
        """..._univs.py:232-303"""
        mf, uv = cfg.MODEL.MASK_FORMER, cfg.MODEL.UniVS
        sampler = None
        if uv.VISUAL_PROMPT_ENCODER:
            sampler = VisualPromptSampler(
                pretrain_img_size=cfg.INPUT.LSJ_AUG.IMAGE_SIZE, hidden_dim=mf.HIDDEN_DIM, num_heads=mf.NHEADS,
                num_frames=cfg.INPUT.SAMPLING_FRAME_NUM, num_prev_frames_memory=uv.TEST.NUM_PREV_FRAMES_MEMORY,
                num_dense_points=uv.VISUAL_PROMPT_PIXELS_PER_IMAGE,
                position_embedding_sin3d_type=uv.POSITION_EMBEDDING_SINE3D, clip_stride=cfg.MODEL.BoxVIS.TEST.CLIP_STRIDE)
        assert mf.DEC_LAYERS >= 1
        return dict(
            in_channels=in_channels, mask_classification=mask_classification,
            num_classes=cfg.MODEL.SEM_SEG_HEAD.NUM_CLASSES, hidden_dim=mf.HIDDEN_DIM,
            num_queries=mf.NUM_OBJECT_QUERIES, nheads=mf.NHEADS, dim_feedforward=mf.DIM_FEEDFORWARD,
            dec_layers=mf.DEC_LAYERS - 1, pre_norm=mf.PRE_NORM, enforce_input_project=mf.ENFORCE_INPUT_PROJ,
            mask_dim=cfg.MODEL.SEM_SEG_HEAD.MASK_DIM, num_frames=cfg.INPUT.SAMPLING_FRAME_NUM,
            clip_class_embed_path=uv.CLIP_CLASS_EMBED_PATH, visual_prompt_sampler=sampler,
            num_dense_points=uv.VISUAL_PROMPT_PIXELS_PER_IMAGE, text_prompt_enable=uv.TEXT_PROMPT_ENCODER,
            prompt_as_queries=uv.PROMPT_AS_QUERIES, text_prompt_to_image_enable=uv.TEXT_PROMPT_TO_IMAGE_ENABLE,
            maskdec_self_attn_mask_type=uv.MASKDEC_SELF_ATTN_MASK_TYPE,
            disable_learnable_queries_sa1b=uv.DISABLE_LEARNABLE_QUERIES_SA1B,
            prompt_self_attn_layers=uv.PROMPT_SELF_ATTN_LAYERS, position_embedding_sin3d_type=uv.POSITION_EMBEDDING_SINE3D,
            num_prev_frames_memory=uv.TEST.NUM_PREV_FRAMES_MEMORY,
            enabled_prev_frames_memory=uv.TEST.ENABLED_PREV_FRAMES_MEMORY,
            enabled_prev_visual_prompts_for_grounding=uv.TEST.ENABLED_PREV_VISUAL_PROMPTS_FOR_GROUNDING,
            semantic_extraction_enable=uv.TEST.SEMANTIC_EXTRACTION.ENABLE)

    def _build(self, in_channels, mask_classification=True, *, num_classes, hidden_dim, num_queries, nheads,
               dim_feedforward, dec_layers, pre_norm, mask_dim, enforce_input_project, prompt_self_attn_layers=-1,
               num_frames=1, clip_class_embed_path, visual_prompt_sampler, num_dense_points, text_prompt_enable=True,
               prompt_as_queries=True, text_prompt_to_image_enable=True, maskdec_self_attn_mask_type="sep",
               disable_learnable_queries_sa1b=False, position_embedding_sin3d_type="FixedT", num_prev_frames_memory=5,
               enabled_prev_frames_memory=True, enabled_prev_visual_prompts_for_grounding=False,
               semantic_extraction_enable=False):
        assert mask_classification, "Only support mask classification model"
        if pre_norm:
            raise NotImplementedError("PRE_NORM=True is not used by any UniVS config (Base.yaml: PRE_NORM False)")
        if hidden_dim != 32 * nheads:
            raise ValueError("the B200 attention kernels require head_dim == 32")
        if in_channels != hidden_dim or enforce_input_project:
            raise NotImplementedError("decoder input_proj convs (ENFORCE_INPUT_PROJ) are not used by any UniVS config")
        if position_embedding_sin3d_type != "ArbitraryT":
            raise NotImplementedError("only POSITION_EMBEDDING_SINE3D='ArbitraryT' (univs/config.py:134 default) is built")
        self.mask_classification = True
        self.num_frames, self.num_heads, self.num_layers = num_frames, nheads, dec_layers
        self.hidden_dim, self.num_queries = hidden_dim, num_queries
        d = hidden_dim
        self.prompt_self_attn_layers = dec_layers if prompt_self_attn_layers < 0 else prompt_self_attn_layers
        self.transformer_self_attention_layers = nn.ModuleList(_SelfAttnLayer(d) for _ in range(dec_layers))
        self.transformer_cross_attention_layers = nn.ModuleList(_CrossAttnLayer(d) for _ in range(dec_layers))
        self.transformer_ffn_layers = nn.ModuleList(_FFNLayer(d, dim_feedforward) for _ in range(dec_layers))
        self.transformer_prompt_self_attention_layers = nn.ModuleList(
            _CrossAttnLayer(d) for _ in range(min(dec_layers, self.prompt_self_attn_layers)))
        self.decoder_norm = nn.LayerNorm(d)
        self.query_feat = nn.Embedding(num_queries, d)
        self.query_embed = nn.Embedding(num_queries, d)
        self.num_feature_levels = 3
        self.level_embed = nn.Embedding(3, d)
        self.input_proj = nn.ModuleList(nn.Sequential() for _ in range(3))
        self.mask_embed = _MLP(d, d, mask_dim, 3)
        # CLIP class embeddings: a plain attribute loaded from disk at construction (..._univs.py:193), not a buffer
        emb = clip_class_embed_path
        self.clip_cls_text_emb = emb if torch.is_tensor(emb) else torch.load(emb, map_location="cpu")
        self.text_emb_dim = self.clip_cls_text_emb.shape[-1]
        self.vis2text_projection = nn.Linear(d, self.text_emb_dim)
        self.text_norm = nn.LayerNorm(self.text_emb_dim)
        self.text2vis_projection = nn.Linear(self.text_emb_dim, d)
        self.cls_temp = nn.Embedding(1, 1)
        self.reid_temp = nn.Embedding(1, 1)
        self.maskdec_self_attn_mask_type = maskdec_self_attn_mask_type
        self.prompt_detection = nn.Embedding(1, d)
        self.prompt_sot = nn.Embedding(1, d)
        self.prompt_grounding = nn.Embedding(1, d)
        self.visual_prompt_sampler = visual_prompt_sampler
        self.num_dense_points = num_dense_points
        self.visual_prompt_enable = visual_prompt_sampler is not None
        self.text_prompt_enable = text_prompt_enable
        self.prompt_as_queries = prompt_as_queries
        self.text_prompt_to_image_enable = text_prompt_to_image_enable
        if text_prompt_to_image_enable:
            self.lang2vision_cross_attention_layer = _CrossAttnLayer(d)
        self.num_prev_frames_memory = max(num_prev_frames_memory, num_frames)
        self.enabled_prev_frames_memory = enabled_prev_frames_memory
        self.enabled_prev_visual_prompts_for_grounding = enabled_prev_visual_prompts_for_grounding
        self.semantic_extraction_enable = semantic_extraction_enable
        self.return_aux_outputs = False
        # diagnostic hook (tests/tools/parity_at_scale.py): fn(call_index, bits, row_open) -> (bits, row_open); lets a parity
        # run record the attention-mask decisions or replay another run's decisions.  None in normal operation.
        self.attn_mask_hook = None
        # intermediate heads from pooled mask features (csrc/decoder_glue.cu): resize(E.F) = E.resize(F); opt-in
        self.pooled_masks = switches.get("POOLED_MASKS") == 1
        self._head_calls = 0
        self._clip_norm_cache = None
        self.eval()

    def _load_from_state_dict(self, state_dict, prefix, *a, **kw):
        # checkpoint upgrade hook of ..._univs.py:32-53: static_query -> query_feat
        for k in list(state_dict.keys()):
            if k.startswith(prefix) and "static_query" in k:
                state_dict[k.replace("static_query", "query_feat")] = state_dict.pop(k)
        return super()._load_from_state_dict(state_dict, prefix, *a, **kw)

    # ------------------------------------------------------------------ helpers
    def _clip_normalized(self, device):
        # keyed on the table's identity and version: `clip_cls_text_emb` is a plain attribute that callers may replace
        # (another vocabulary file) or edit in place after the first forward
        if self.clip_cls_text_emb.device != device:
            self.clip_cls_text_emb = self.clip_cls_text_emb.to(device)
        t = self.clip_cls_text_emb
        key = (t.data_ptr(), t._version, tuple(t.shape), str(device))
        c = self._clip_norm_cache
        if c is None or c[0] != key or c[2]() is not t:
            import weakref
            c = (key, F.normalize(t.float(), p=2, dim=-1), weakref.ref(t))
            self._clip_norm_cache = c
        return c[1]

    def _self_attn_mask_bits(self, t, n_lp, device, task):
        """generate_self_attn_mask (:824-848) over tokens ordered (q*T + t); returns packed bits or None when
        nothing is blocked."""
        kind = self.maskdec_self_attn_mask_type
        if kind in ("none", "all"):
            return None
        nq = self.num_queries
        np_ = n_lp - nq
        if np_ == 0:
            return None                                      # learnable block only -> all visible
        m = torch.ones((n_lp * t, n_lp * t), dtype=torch.bool, device=device)
        m[: nq * t, : nq * t] = False
        if kind == "sep-blocked" or task == "grounding":
            blk = torch.block_diag(*[torch.ones(t, t, dtype=torch.bool, device=device)] * np_)
            m[nq * t:, nq * t:] = ~blk
        elif kind == "sep":
            m[nq * t:, nq * t:] = False
        elif kind == "sep-l2p":
            m[nq * t:] = False
        else:
            raise ValueError(kind)
        return ops.pack_mask_bits(m[None])

    def _cross_attention(self, layer, x, qpos, k, v, bits, row_open):
        mha = layer.multihead_attn
        wq, bq = mha.wq()
        q = nn_ops.linear(x + qpos, wq, bq)
        a = ops.mha_core(q, k, v, bits, row_open)
        o = nn_ops.linear(a, mha.out_proj.weight, None)
        # k, v arrive without their in-proj biases: the key bias shifts every score of a row by the same constant
        # (softmax-invariant) and the value bias is folded into the output bias
        return nn_ops.layernorm(x, layer.norm, residual=o, for_gemm=False, residual_bias=mha.folded_out_bias())[1]

    def _self_attention(self, layer, x, qpos, bits):
        """x, qpos: [T,Q,C] -> tokens (q*T + t) (..._univs.py:408-416)"""
        T, Q, C = x.shape
        mha = layer.self_attn
        xs = x.transpose(0, 1).reshape(1, Q * T, C)
        ps = qpos.transpose(0, 1).reshape(1, Q * T, C)
        wqk, bqk = mha.wqk()
        qk = nn_ops.linear(xs + ps, wqk, bqk)
        wv, bv = mha.wv()
        v = nn_ops.linear(xs, wv, bv)
        a = ops.mha_core(qk[..., :C].contiguous(), qk[..., C:].contiguous(), v, bits, None)
        o = nn_ops.linear(a, mha.out_proj.weight, None)
        y = nn_ops.layernorm(xs.contiguous(), layer.norm, residual=o, for_gemm=False, residual_bias=mha.out_proj.bias)[1]
        return y.view(Q, T, C).transpose(0, 1).contiguous()

    def _proca(self, i, x, qpos, mem, mem_pe):
        """ProCA (:456-496).  x,qpos [T,Q,C]; mem [P,Tm,L,C] prompt memory (+ mem_pe or None)."""
        nq = self.num_queries
        if x.shape[1] == nq:
            return x
        layer = self.transformer_prompt_self_attention_layers[i]
        mha = layer.multihead_attn
        tok = x[:, nq:].transpose(0, 1).contiguous()                      # [P,T,C]
        wq, bq = mha.wq(); wk, bk = mha.wk(); wv, bv = mha.wv()
        lin = nn_ops.linear_prepped
        tk = nn_ops.prep(tok)
        mm = nn_ops.prep(mem)
        if mem_pe is not None:
            tq = nn_ops.prep(tok + qpos[:, nq:].transpose(0, 1))
            q, k_self = lin(tq, wq, bq), lin(tq, wk, bk)
            k_mem = nn_ops.linear(mem + mem_pe, wk, bk)
        else:
            q, k_self = lin(tk, wq, bq), lin(tk, wk, bk)
            k_mem = lin(mm, wk, bk)
        v_self = lin(tk, wv, bv)
        v_mem = lin(mm, wv, bv)
        a = ops.proca_core(q.contiguous(), k_self.contiguous(), v_self.contiguous(), k_mem.contiguous(), v_mem.contiguous())
        o = nn_ops.linear(a, mha.out_proj.weight, None)
        y = nn_ops.layernorm(tok, layer.norm, residual=o, for_gemm=False, residual_bias=mha.out_proj.bias)[1]   # [P,T,C]
        return torch.cat([x[:, :nq], y.transpose(0, 1)], 1)

    def _heads(self, x, feats_cl, hw, next_hw, task, targets, t, need_class, need_attn, out_buf=None, exchange=None,
               pooled_feats=None):
        """forward_prediction_heads (:498-567).  x [T,Q,C] -> (class logits | None, mask logits [Q,T,HW], bits, row_open, reid)"""
        dec, dec_g = nn_ops.layernorm(x, self.decoder_norm, want_sum=False, for_gemm=False)[1], None
        dec_g = nn_ops.prep(dec)                                            # GEMM operand of decoder_norm(x)
        cls, reid = None, [None]
        if need_class or (task == "grounding" and self.prompt_as_queries):
            oc = nn_ops.linear_prepped(dec_g, self.vis2text_projection.weight, self.vis2text_projection.bias)   # [T,Q,640]
            if task != "grounding":
                if need_class:
                    clip = self._clip_normalized(x.device)
                    # mean over T commutes with the (linear) class einsum: average first, 1/T of the work
                    ocn = F.normalize(oc, p=2, dim=-1)
                    if exchange is not None:                                # frame-sharded decoder: the mean needs all frames
                        ocn = exchange.gather(ocn)
                    cls = nn_ops.linear(ocn.mean(0, keepdim=True), clip) * self.cls_temp.weight.exp()
            else:
                exp = torch.stack([tg["exp_sentence_feats"][:, 0] for tg in targets]).to(oc)      # [1,P,640]
                cls = nn_ops.linear(oc.mean(0, keepdim=True), exp[0].contiguous(), cache=False)
        emb = self.mask_embed(dec_g, prepped=True)                          # [T,Q,C]
        if pooled_feats is not None:
            # intermediate head: its mask logits only feed the next layer's attention mask, and the bilinear resize to the
            # memory size commutes with the einsum -- run it on the pooled features, at the memory resolution
            small = ops.mask_einsum(emb.contiguous(), pooled_feats, tag="mask_einsum_pooled")      # [Q,T,S_next]
            bits, row_open = ops.attn_mask_bits_direct(small)
            if self.attn_mask_hook is not None:
                bits, row_open = self.attn_mask_hook(self._head_calls, bits, row_open)
            self._head_calls += 1
            return cls, None, bits, row_open, reid
        logits = ops.mask_einsum(emb.contiguous(), feats_cl, out=out_buf)   # [Q,T,HW]
        if task == "grounding" and self.prompt_as_queries:
            # learnable-for-prompt mask fusion (:537-547)
            nq = self.num_queries
            on = F.normalize(dec, p=2, dim=-1)
            with nn_ops.ieee_fp32():
                reid = torch.einsum("tqc,tkc->tqk", on, on[:, nq:]).mean(0, keepdim=True)  # [1,Q,P]
            idx = reid[0, :nq].argmax(0)
            logits[nq:] = (logits[nq:] + logits[idx]) / 2.0
        bits = row_open = None
        if need_attn:
            bits, row_open = ops.attn_mask_bits(logits, hw, next_hw)
            if self.attn_mask_hook is not None:
                bits, row_open = self.attn_mask_hook(self._head_calls, bits, row_open)
        self._head_calls += 1
        return cls, logits, bits, row_open, reid

    # ------------------------------------------------------------------ prompts
    def _lang_to_vision(self, feats, src):
        """forward_lang_to_vision (:760-793), inference part: text tokens attend to all 3 levels.
        feats [T, Np, C]; src list of [T,S_l,C].  The averaged attention weights the reference also returns
        (need_weights=True) are only used by a training loss and are not computed."""
        layer = self.lang2vision_cross_attention_layer
        mha = layer.multihead_attn
        mem = torch.cat(src, 1)
        wq, bq = mha.wq(); wk, bk = mha.wk(); wv, bv = mha.wv()
        mm = nn_ops.prep(mem)
        a = ops.mha_core(nn_ops.linear(feats, wq, bq), nn_ops.linear_prepped(mm, wk, bk), nn_ops.linear_prepped(mm, wv, bv))
        o = nn_ops.linear(a, mha.out_proj.weight, None)
        return nn_ops.layernorm(feats.contiguous(), layer.norm, residual=o, for_gemm=False, residual_bias=mha.out_proj.bias)[1]

    def _prompt_encoder(self, src, pos, size_list, targets, t):
        """forward_prompt_encoder (:599-758), inference branches.
        Returns (prompt tokens [T,P,C] | None, prompt query-pe [T,P,C] | None, memory [P,Tm,L,C] | None,
        memory pe | None)."""
        task = targets[0]["task"]
        device = src[0].device
        if task == "sot" or targets[0]["prompt_type"] == "visual":
            if self.visual_prompt_sampler is None:
                raise AttributeError("PROMPT_AS_QUERIES with prompt_type='visual' needs VISUAL_PROMPT_ENCODER=True "
                                     "(the reference dereferences a None sampler here, ..._univs.py:631-634)")
            pe_dense, feats_dense = self.visual_prompt_sampler.process_per_batch(src, pos, size_list, targets)
            if feats_dense is None:
                return None, None, None, None
            # feats_dense / pe_dense: [P, L, T, C]; blank (all-zero) prompts are excluded from the mean (:636-643)
            nb_f = (~(feats_dense == 0).all(-1)).unsqueeze(-1).sum(1).clamp(min=1)
            nb_p = (~(pe_dense == 0).all(-1)).unsqueeze(-1).sum(1).clamp(min=1)
            f_mean = feats_dense.sum(1) / nb_f                               # [P,T,C]
            p_mean = pe_dense.sum(1) / nb_p
            tokens = f_mean + self.prompt_sot.weight.view(1, 1, -1)
            if "prompt_feats" in targets[0]:
                pe_dense, feats_dense = self.visual_prompt_sampler.memory_pool_prompts(
                    targets[0], self.num_prev_frames_memory)                  # [P,1,L',C] (T-invariant)
            else:
                feats_dense = feats_dense.permute(0, 2, 1, 3).contiguous()    # [P,T,L,C]
                pe_dense = pe_dense.permute(0, 2, 1, 3).contiguous()
            return tokens.transpose(0, 1).contiguous(), p_mean.transpose(0, 1).contiguous(), feats_dense, pe_dense
        if task == "detection":
            name = targets[0]["dataset_name"]
            assert name in COMBINED_DATASETS_CATEGORY_INFO
            n_cls, start = COMBINED_DATASETS_CATEGORY_INFO[name]
            emb = self.clip_cls_text_emb.to(device)[start:start + n_cls].float()
            assert len(emb) == n_cls, f"Dismatch numbers of class, {len(emb)} and {n_cls}"
            f = nn_ops.linear_prepped(nn_ops.layernorm(emb, self.text_norm)[1], self.text2vis_projection.weight,
                                      self.text2vis_projection.bias)            # [P,C]
            feats = f[None].expand(t, -1, -1).contiguous()                  # [T,P,C]
            if self.text_prompt_to_image_enable:
                feats = self._lang_to_vision(feats, src)
                mem = feats.transpose(0, 1).unsqueeze(2).contiguous()       # [P,T,1,C]
            else:
                mem = f[:, None, None, :].contiguous()                      # [P,1,1,C] (T-invariant)
            return feats + self.prompt_detection.weight.view(1, 1, -1), feats, mem, None
        if task == "grounding":
            tg = targets[0]
            words = tg["exp_word_feats"][..., :t, :].to(device)             # [P,77,T,640]
            sent = tg["exp_sentence_feats"][..., :t, :].to(device)          # [P,T,640]
            P, Lw = words.shape[:2]
            ef = torch.cat([sent[:, None], words], 1)                       # [P,78,T,640]
            f = nn_ops.linear_prepped(nn_ops.layernorm(ef.float(), self.text_norm)[1], self.text2vis_projection.weight,
                                      self.text2vis_projection.bias)            # [P,78,T,C]
            feats = f.permute(2, 0, 1, 3).reshape(t, P * (Lw + 1), -1).contiguous()   # [T, P*78, C]
            if self.text_prompt_to_image_enable:
                feats = self._lang_to_vision(feats, src)
            dense = feats.view(t, P, Lw + 1, -1)
            sentence = dense[:, :, 0]                                       # [T,P,C]
            mem = dense.permute(1, 0, 2, 3).contiguous()                    # [P,T,78,C]
            return sentence + self.prompt_grounding.weight.view(1, 1, -1), sentence.contiguous(), mem, None
        raise ValueError(task)

    # ------------------------------------------------------------------ forward
    def supports_exchange(self, targets):
        """The frame-sharded (token-exchange) decoder covers the clip paths whose prompt handling is per-frame:
        detection with learnable queries and with category prompts.  Visual prompts (memory pool across frames) and
        grounding (cross-frame reid fusion) go through the feature all-gather instead."""
        tg = targets[0]
        has_masks = "masks" in tg and torch.is_tensor(tg["masks"]) and tg["masks"].nelement() > 0
        return (tg["task"] == "detection" and not has_masks and "prompt_feats" not in tg
                and not self.return_aux_outputs and not self.semantic_extraction_enable)

    @torch.no_grad()
    def forward(self, x, mask_features, mask_features_bfe_conv=None, mask=None, targets=None, exchange=None):
        """exchange: None, or a sharding.TokenExchange -- x / mask_features then hold only this rank's frames and the
        decoder stays frame-sharded (SURVEY.md 8e "cheaper alternative"); every rank returns the full-clip outputs."""
        if self.training:
            raise NotImplementedError("the B200 decoder implements the inference path only")
        assert len(x) == self.num_feature_levels
        t, c_m, h_m, w_m = mask_features.shape                              # bs = 1 at inference (:309-311)
        device = mask_features.device
        task = targets[0]["task"]
        t_all = t
        if exchange is not None:
            if not self.supports_exchange(targets):
                raise NotImplementedError("token-exchange decoder: detection clips without visual prompts only")
            t_all = exchange.num_frames
            assert t == len(exchange.frames)
        feats_cl = mask_features.permute(0, 2, 3, 1)                        # [T,H,W,C]
        if not feats_cl.is_contiguous():
            feats_cl = feats_cl.contiguous()
        feats_raw = feats_cl.view(t, h_m * w_m, c_m)
        feats_cl = ops.prepare_mask_features(feats_raw)
        # the temporal term of the position encoding is the same for the three levels: once per clip (cached for host-side
        # or default frame indices)
        pz = position.temporal_sine(targets[0].get("frame_indices", t_all), device, self.level_embed.weight.shape[1] // 2)
        if exchange is not None:
            pz = pz.index_select(0, exchange.frames_tensor(device))
        src, pos, size_list = [], [], []
        for i in range(3):
            n, c, h, w = x[i].shape
            size_list.append((h, w))
            xi = x[i].permute(0, 2, 3, 1).reshape(n, h * w, c)              # token-major (view when channel-last)
            src.append(xi + self.level_embed.weight[i])
            pos.append(position.sine_3d_arbitrary_t(None, h, w, device, c // 2, pz=pz))
        nq = self.num_queries
        out = self.query_feat.weight[None].expand(t, -1, -1).contiguous()   # [T,Q,C]
        qpos = self.query_embed.weight[None].expand(t, -1, -1).contiguous()
        mem = mem_pe = None
        if self.prompt_as_queries:
            p_tok, p_pe, mem, mem_pe = self._prompt_encoder(src, pos, size_list, targets, t)
            if p_tok is not None:
                out = torch.cat([out, p_tok], 1)
                qpos = torch.cat([qpos, p_pe if p_pe is not None else p_tok], 1)
            out = self._proca(0, out, qpos, mem, mem_pe)
            qpos = torch.cat([qpos[:, :nq], out[:, nq:]], 1)                # (:366)
        n_lp = out.shape[1]
        aux = []
        want_aux = self.return_aux_outputs

        def record(cls, logits, reid, emb):
            aux.append({"pred_logits": cls, "pred_masks": logits.view(1, n_lp, t, h_m, w_m).clone(),
                        "pred_reid_logits": reid,
                        "pred_embds": nn_ops.layernorm(emb.transpose(0, 1), self.decoder_norm, for_gemm=False)[1][None]})

        hw = (h_m, w_m)
        self._head_calls = 0
        pooled = {}
        if (self.pooled_masks and not want_aux and not (task == "grounding" and self.prompt_as_queries)
                and all(h_m % h == 0 and w_m % w == 0 and (h_m // h) % 2 == 0 and (w_m // w) % 2 == 0 for h, w in size_list)
                # the register (mma.sync) einsum of the non-fp16x3 policies needs an even pixel count per frame
                and (ops._einsum_mode == "f16x3" or all((h * w) % 2 == 0 for h, w in size_list))):
            for size in size_list:         # once per clip: the mask features at the three memory resolutions
                if size not in pooled:
                    pooled[size] = ops.mask_feature_pool(feats_raw, hw, size)
        cls, logits, bits, row_open, reid = self._heads(out, feats_cl, hw, size_list[0], task, targets, t, want_aux, True,
                                                        exchange=exchange, pooled_feats=pooled.get(size_list[0]))
        if want_aux:
            record(cls, logits, reid, out)
        sa_bits = self._self_attn_mask_bits(t_all, n_lp, device, task)
        qpos_all = exchange.gather(qpos) if exchange is not None else None   # constant over the layers
        kv_operands = {}
        for i in range(self.num_layers):
            if self.prompt_as_queries and 0 < i < self.prompt_self_attn_layers:
                out = self._proca(i, out, qpos, mem, mem_pe)
            lvl = i % 3
            ca = self.transformer_cross_attention_layers[i].multihead_attn
            wk, bk = ca.wk(); wv, bv = ca.wv()
            if lvl not in kv_operands:      # every level feeds three layers: its GEMM operands are prepared once
                kv_operands[lvl] = (nn_ops.prep(src[lvl] + pos[lvl]), nn_ops.prep(src[lvl]))
            k = nn_ops.linear_prepped(kv_operands[lvl][0], wk, None)      # biases handled in _cross_attention
            v = nn_ops.linear_prepped(kv_operands[lvl][1], wv, None)
            out = self._cross_attention(self.transformer_cross_attention_layers[i], out, qpos, k, v, bits, row_open)
            if exchange is None:
                out = self._self_attention(self.transformer_self_attention_layers[i], out, qpos, sa_bits)
            else:       # the Q*T self-attention is the one step that sees all frames: tokens travel, features stay
                out = exchange.local(self._self_attention(self.transformer_self_attention_layers[i],
                                                          exchange.gather(out), qpos_all, sa_bits))
            out = self.transformer_ffn_layers[i](out)
            last = i == self.num_layers - 1
            cls, logits, bits, row_open, reid = self._heads(
                out, feats_cl, hw, size_list[(i + 1) % 3], task, targets, t, want_aux or last, not last, out_buf=logits,
                exchange=exchange, pooled_feats=None if last else pooled.get(size_list[(i + 1) % 3]))
            if want_aux and not last:
                record(cls, logits, reid, out)
        embds = nn_ops.layernorm(out.transpose(0, 1), self.decoder_norm, for_gemm=False)[1][None]   # [1,Q,T,C]
        if exchange is not None:        # full-clip outputs on every rank
            embds = exchange.gather(embds[0].transpose(0, 1)).transpose(0, 1).contiguous()[None]
            logits = exchange.gather(logits.transpose(0, 1)).transpose(0, 1).contiguous()
        result = {
            "pred_logits": cls,
            "pred_masks": logits.view(1, n_lp, t_all, h_m, w_m),
            "aux_outputs": aux,
            "pred_embds": embds,
            "pred_reid_logits": reid,
        }
        if self.semantic_extraction_enable:
            result.update({"pred_embds": out.transpose(0, 1).permute(1, 2, 0), "mask_features": mask_features})
        return result
