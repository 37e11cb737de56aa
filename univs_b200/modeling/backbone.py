"""Swin backbone, B200-native host side.

Drop-in for mask2former/modeling/backbone/swin.py (`SwinTransformer` :498-683, `D2SwinTransformer` :686-772): same
constructor arguments, same `state_dict` keys/shapes, same forward contract
`x [N,3,H,W] -> {"res2".."res5": [N, E*2^i, H/2^(i+2), W/2^(i+2)]}`.

What is different underneath: activations stay token-major / channel-last ([N,H,W,C]) for the whole backbone; the
window partition / cyclic shift / padding / reverse copies of swin.py:247-289 do not exist -- the fused
`swin_window_attention` kernel reads the qkv grid through that addressing and writes the result back through its
inverse; the relative-position index buffer is kept only for state_dict parity (the kernel computes it).
The returned NCHW feature maps are zero-copy views of channel-last storage.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import nn_ops, ops
from ..registry import BACKBONE_REGISTRY, ShapeSpec


def _relative_position_index(ws: int) -> torch.Tensor:
    ar = torch.arange(ws)
    cy, cx = torch.meshgrid(ar, ar, indexing="ij")
    cy, cx = cy.reshape(-1), cx.reshape(-1)
    return (cy[:, None] - cy[None, :] + ws - 1) * (2 * ws - 1) + (cx[:, None] - cx[None, :] + ws - 1)


class _WindowAttentionParams(nn.Module):
    """Parameter holder with the key names of swin.py:74-129 (`attn.*`)."""

    def __init__(self, dim, window, heads, qkv_bias=True):
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window - 1) ** 2, heads))
        self.register_buffer("relative_position_index", _relative_position_index(window))
        self.qkv = nn.Linear(dim, 3 * dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, heads, window, shift, mlp_ratio, qkv_bias):
        super().__init__()
        self.heads, self.window, self.shift = heads, window, shift
        self.norm1 = nn.LayerNorm(dim)
        self.attn = _WindowAttentionParams(dim, window, heads, qkv_bias)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x, pending, pending_bias=None):
        """x: residual stream [N,H,W,C]; pending (+ pending_bias): branch output (and the deferred bias of the GEMM that
        produced it) still to be added to the stream.  Residual adds and GEMM biases are folded into the fused kernels
        that consume them (LayerNorm, GELU, window attention); LayerNorm / GELU outputs are emitted directly in the GEMM
        operand format of the active policy (nn_ops)."""
        a = self.attn
        if nn_ops.gemm_tc() and a.qkv.weight.shape[1] <= 1536 and x.is_contiguous():
            return self._forward_fused_residual(x, pending, pending_bias)
        x, h = nn_ops.layernorm(x, self.norm1, residual=pending, want_sum=True, residual_bias=pending_bias)
        qkv = nn_ops.linear_prepped(h, a.qkv.weight, None)                  # bias added inside the attention kernel
        bias = a.qkv.bias if a.qkv.bias is not None else x.new_zeros(qkv.shape[-1])
        if nn_ops.policy() == "fp16x3" and qkv.shape[-1] <= 3 * 1536:     # attention writes the GEMM operand itself
            ys = ops.swin_window_attention_operand(qkv, bias, a.relative_position_bias_table, self.heads, self.window,
                                                   self.shift, compact=nn_ops.gemm_tc())
            y = nn_ops.linear_prepped(ys, a.proj.weight, None)
        else:
            y = ops.swin_window_attention(qkv, bias, a.relative_position_bias_table, self.heads, self.window, self.shift)
            y = nn_ops.linear(y, a.proj.weight, None)
        x, h = nn_ops.layernorm(x, self.norm2, residual=y, want_sum=True, residual_bias=a.proj.bias)
        z = nn_ops.mlp(h, self.mlp.fc1, self.mlp.fc2)      # fc1 -> GELU (+ deferred fc1 bias) -> fc2 (optionally in L2-sized row chunks)
        return x, z, self.mlp.fc2.bias


    def _forward_fused_residual(self, x, pending, pending_bias):
        """Own-GEMM variant: both residual adds (and the proj / fc2 biases) ride in the GEMM epilogues, which update the
        residual stream in place; the LayerNorms read one tensor and write only the next operand."""
        a = self.attn
        if pending is not None:                    # first block after a library-path producer
            x, h = nn_ops.layernorm(x, self.norm1, residual=pending, want_sum=True, residual_bias=pending_bias)
        else:
            h = nn_ops.layernorm(x, self.norm1)[1]
        qkv = nn_ops.linear_prepped(h, a.qkv.weight, None)
        bias = a.qkv.bias if a.qkv.bias is not None else x.new_zeros(qkv.shape[-1])
        ys = ops.swin_window_attention_operand(qkv, bias, a.relative_position_bias_table, self.heads, self.window, self.shift,
                                               compact=True)
        x = nn_ops.linear_residual(ys, a.proj.weight, a.proj.bias, x)
        h = nn_ops.layernorm(x, self.norm2)[1]
        x = nn_ops.mlp(h, self.mlp.fc1, self.mlp.fc2, residual=x)
        return x, None, None


class _PatchMerging(nn.Module):
    """swin.py:298-337"""

    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)

    def forward(self, x):                      # [N,H,W,C] -> [N,ceil(H/2),ceil(W/2),2C]
        N, H, W, C = x.shape
        if nn_ops.fused_glue() and 4 * C <= 4096:       # gather + LayerNorm in one kernel, no concatenated copy
            return nn_ops.linear_prepped(nn_ops.layernorm_merge2x2(x, self.norm), self.reduction.weight, None)
        if H % 2 or W % 2:
            x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
        x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
        _, h = nn_ops.layernorm(x, self.norm)
        return nn_ops.linear_prepped(h, self.reduction.weight, None)


class _Stage(nn.Module):
    def __init__(self, dim, depth, heads, window, mlp_ratio, qkv_bias, downsample):
        super().__init__()
        self.blocks = nn.ModuleList(
            _Block(dim, heads, window, 0 if i % 2 == 0 else window // 2, mlp_ratio, qkv_bias) for i in range(depth))
        self.downsample = _PatchMerging(dim) if downsample else None

    def forward(self, x, out_norm=None):
        """-> (norm_i(x_out) or None, x_out or downsample(x_out))"""
        pending = pbias = None
        for blk in self.blocks:
            x, pending, pbias = blk(x, pending, pbias)
        y = None
        if out_norm is not None and pending is None:        # residuals already folded in (own-GEMM epilogues)
            y = nn_ops.layernorm(x, out_norm, for_gemm=False)[1]
        elif out_norm is not None:
            x, y = nn_ops.layernorm(x, out_norm, residual=pending, want_sum=True, for_gemm=False, residual_bias=pbias)
        elif pending is not None:
            x = x + pending + pbias
        return y, (self.downsample(x) if self.downsample is not None else x)


class _PatchEmbed(nn.Module):
    """swin.py:456-495"""

    def __init__(self, patch, in_chans, dim, norm):
        super().__init__()
        self.patch = patch
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)
        self.norm = nn.LayerNorm(dim) if norm else None

    def forward(self, x):                      # [N,3,H,W] -> [N,H/4,W/4,C]
        p = self.patch
        _, _, H, W = x.shape
        if W % p or H % p:
            x = F.pad(x, (0, (p - W % p) % p, 0, (p - H % p) % p))
        with nn_ops.ieee_fp32():                      # 3->E 4x4 patch projection: tiny, kept in exact fp32
            x = self.proj(x).permute(0, 2, 3, 1).contiguous()
        return nn_ops.layernorm(x, self.norm, for_gemm=False)[1] if self.norm is not None else x

    def forward_frames(self, frames, pixel_mean, pixel_std, padded_size):
        """Fused ingest (csrc/swin_glue.cu): raw frames [N,3,H,W] uint8 / float32 -> normalise, zero-pad to
        `padded_size` and gather 4x4 patches in one pass; the patch projection is then one GEMM (weight.view(E, 48))."""
        p = self.patch
        Hp, Wp = padded_size
        padded = ((Hp + p - 1) // p * p, (Wp + p - 1) // p * p)
        rows = nn_ops.patchify(frames, pixel_mean, pixel_std, padded, p)
        x = nn_ops.linear_prepped(rows, self.proj.weight.view(self.proj.out_channels, -1), self.proj.bias)
        return nn_ops.layernorm(x, self.norm, for_gemm=False)[1] if self.norm is not None else x


class SwinTransformer(nn.Module):
    def __init__(self, pretrain_img_size=224, patch_size=4, in_chans=3, embed_dim=96, depths=(2, 2, 6, 2),
                 num_heads=(3, 6, 12, 24), window_size=7, mlp_ratio=4.0, qkv_bias=True, qk_scale=None, drop_rate=0.0,
                 attn_drop_rate=0.0, drop_path_rate=0.2, norm_layer=nn.LayerNorm, ape=False, patch_norm=True,
                 out_indices=(0, 1, 2, 3), frozen_stages=-1, use_checkpoint=False):
        super().__init__()
        if ape:
            raise NotImplementedError("absolute position embedding (MODEL.SWIN.APE) is not used by any UniVS config")
        if qk_scale is not None:
            raise NotImplementedError("qk_scale override is not supported (head_dim**-0.5 is built into the kernel)")
        for d, h in zip([embed_dim * 2 ** i for i in range(len(depths))], num_heads):
            if d != 32 * h:
                raise ValueError("the B200 window-attention kernel requires head_dim == 32 (true for Swin-T/S/B/L)")
        self.embed_dim, self.out_indices, self.num_layers = embed_dim, tuple(out_indices), len(depths)
        self.patch_embed = _PatchEmbed(patch_size, in_chans, embed_dim, patch_norm)
        self.layers = nn.ModuleList(
            _Stage(embed_dim * 2 ** i, depths[i], num_heads[i], window_size, mlp_ratio, qkv_bias,
                   downsample=i < len(depths) - 1) for i in range(len(depths)))
        self.num_features = [embed_dim * 2 ** i for i in range(len(depths))]
        for i in self.out_indices:
            self.add_module(f"norm{i}", nn.LayerNorm(self.num_features[i]))
        self.eval()

    @torch.no_grad()
    def forward_frames(self, frames, pixel_mean, pixel_std, padded_size):
        """Same as forward((frames - mean) / std zero-padded to padded_size) without materialising that tensor."""
        if self.patch_embed.patch != 4:
            raise NotImplementedError("fused ingest is built for the 4x4 patch embedding")
        return self._stages(self.patch_embed.forward_frames(frames, pixel_mean, pixel_std, padded_size))

    @torch.no_grad()
    def forward(self, x):
        return self._stages(self.patch_embed(x))

    def _stages(self, x):
        outs = {}
        for i, stage in enumerate(self.layers):
            y, x = stage(x, getattr(self, f"norm{i}") if i in self.out_indices else None)
            if y is not None:                                   # [N,H,W,C] contiguous
                outs[f"res{i + 2}"] = y.permute(0, 3, 1, 2)    # NCHW view of channel-last storage
        return outs


@BACKBONE_REGISTRY.register()
class D2SwinTransformer(SwinTransformer):
    """swin.py:686-772: built from cfg.MODEL.SWIN.*"""

    def __init__(self, cfg, input_shape=None):
        s = cfg.MODEL.SWIN
        super().__init__(s.PRETRAIN_IMG_SIZE, s.PATCH_SIZE, 3, s.EMBED_DIM, s.DEPTHS, s.NUM_HEADS, s.WINDOW_SIZE,
                         s.MLP_RATIO, s.QKV_BIAS, s.QK_SCALE, s.DROP_RATE, s.ATTN_DROP_RATE, s.DROP_PATH_RATE,
                         nn.LayerNorm, s.APE, s.PATCH_NORM, frozen_stages=cfg.MODEL.BACKBONE.FREEZE_AT,
                         use_checkpoint=s.USE_CHECKPOINT)
        self._out_features = list(s.OUT_FEATURES)
        self._out_feature_strides = {"res2": 4, "res3": 8, "res4": 16, "res5": 32}
        self._out_feature_channels = {f"res{i + 2}": self.num_features[i] for i in range(4)}

    def forward(self, x):
        if x.dim() != 4:
            raise AssertionError(f"SwinTransformer takes an input of shape (N, C, H, W). Got {x.shape} instead!")
        y = super().forward(x)
        return {k: v for k, v in y.items() if k in self._out_features}

    def output_shape(self):
        return {n: ShapeSpec(channels=self._out_feature_channels[n], stride=self._out_feature_strides[n])
                for n in self._out_features}

    @property
    def size_divisibility(self):
        return 32
