"""MSDeformAttn pixel decoder, B200-native host side.

Drop-in for mask2former/modeling/pixel_decoder/msdeformattn.py::MSDeformAttnPixelDecoder (:166-360): same ctor
kwargs / from_config, same state_dict keys, `forward_features(features)` returns
`(mask_features, out[-1], out[0], multi_scale_features)` with the reference's NCHW shapes.

Underneath: the 6 encoder layers run on one token-major tensor [N, Len, 256]; per layer the `sampling_offsets` and
`attention_weights` linears are one fused 256->288 GEMM whose raw output feeds `ms_deform_attn_encoder` (softmax over
L*P, reference-point and sampling-location arithmetic fused into the gather kernel -- no loc/weight tensors);
the FPN level runs channels-last so that `mask_features` is produced channel-last ([N,HW,C] storage) -- the layout
the mask einsum consumes -- and is returned as an NCHW *view* of that storage.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import nn_ops, ops
from ..registry import SEM_SEG_HEADS_REGISTRY, is_cfg
from . import position


class _MSDeformAttnParams(nn.Module):
    """Key names of ops/modules/ms_deform_attn.py:34-80; init per :63-80."""

    def __init__(self, d_model, n_levels, n_heads, n_points):
        super().__init__()
        self.n_levels, self.n_heads, self.n_points = n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        nn.init.zeros_(self.sampling_offsets.weight)
        th = torch.arange(n_heads, dtype=torch.float32) * (2.0 * np.pi / n_heads)
        grid = torch.stack([th.cos(), th.sin()], -1)
        grid = (grid / grid.abs().max(-1, keepdim=True)[0]).view(n_heads, 1, 1, 2).repeat(1, n_levels, n_points, 1)
        for i in range(n_points):
            grid[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias.copy_(grid.view(-1))
        nn.init.zeros_(self.attention_weights.weight)
        nn.init.zeros_(self.attention_weights.bias)
        nn.init.xavier_uniform_(self.value_proj.weight)
        nn.init.zeros_(self.value_proj.bias)
        nn.init.xavier_uniform_(self.output_proj.weight)
        nn.init.zeros_(self.output_proj.bias)
        self._fused = None

    def fused_offs_logits(self):
        """[sampling_offsets ; attention_weights] as one (288 x 256) GEMM; rebuilt if parameters were replaced."""
        key = (self.sampling_offsets.weight.data_ptr(), self.attention_weights.weight.data_ptr(),
               self.sampling_offsets.weight._version, self.attention_weights.weight._version)
        if self._fused is None or self._fused[0] != key:
            w = torch.cat([self.sampling_offsets.weight, self.attention_weights.weight], 0).contiguous()
            b = torch.cat([self.sampling_offsets.bias, self.attention_weights.bias], 0).contiguous()
            self._fused = (key, w, b)
        return self._fused[1], self._fused[2]


class _EncoderLayer(nn.Module):
    def __init__(self, d_model, d_ffn, n_levels, n_heads, n_points):
        super().__init__()
        self.self_attn = _MSDeformAttnParams(d_model, n_levels, n_heads, n_points)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.norm2 = nn.LayerNorm(d_model)

    def forward(self, src, pos, shapes, starts, carry=None, emit_next=False):
        """-> (src_out, carry_out).  `carry` = (operand(src), operand(src + pos)) emitted by the previous layer's last
        LayerNorm under the fused-glue path (None otherwise / for the first layer)."""
        a = self.self_attn
        N, S, C = src.shape
        w, b = a.fused_offs_logits()
        if nn_ops.fused_glue():
            # the three linear biases move into the MSDeformAttn kernel (no bias-broadcast copies for the GEMMs), the kernel
            # emits the operand of output_proj, and each LayerNorm emits the operands of the GEMMs that consume it
            h_src, h_q = carry if carry is not None else (nn_ops.prep(src), nn_ops.prep(src + pos))
            value = nn_ops.linear_prepped(h_src, a.value_proj.weight, None).view(N, S, a.n_heads, C // a.n_heads)
            offs_logits = nn_ops.linear_prepped(h_q, w, None)
            y = ops.ms_deform_attn_encoder(value, shapes, starts, offs_logits, a.n_levels, a.n_points,
                                           value_bias=a.value_proj.bias, offs_logits_bias=b,
                                           split=nn_ops._fmt() if nn_ops.splitting() else None)
            o = nn_ops.linear_prepped(y, a.output_proj.weight, None)
            src, h1, _ = nn_ops.layernorm_multi(src, self.norm1, residual=o, residual_bias=a.output_proj.bias)
            z = nn_ops.linear_prepped(nn_ops.linear_act_operand(h1, self.linear1), self.linear2.weight, None)
            src, h2, hq2 = nn_ops.layernorm_multi(src, self.norm2, residual=z, residual_bias=self.linear2.bias,
                                                  pos=pos if emit_next else None, want_operand=emit_next)
            return src, ((h2, hq2) if emit_next else None)
        value = nn_ops.linear(src, a.value_proj.weight, a.value_proj.bias).view(N, S, a.n_heads, C // a.n_heads)
        offs_logits = nn_ops.linear(src + pos, w, b)
        y = ops.ms_deform_attn_encoder(value, shapes, starts, offs_logits, a.n_levels, a.n_points)
        o = nn_ops.linear(y, a.output_proj.weight, None)
        src = nn_ops.layernorm(src, self.norm1, residual=o, for_gemm=False, residual_bias=a.output_proj.bias)[1]
        z = nn_ops.linear_prepped(nn_ops.linear_act_operand(nn_ops.prep(src), self.linear1), self.linear2.weight, None)
        return nn_ops.layernorm(src, self.norm2, residual=z, for_gemm=False, residual_bias=self.linear2.bias)[1], None


class _Encoder(nn.Module):
    def __init__(self, layers):
        super().__init__()
        self.layers = nn.ModuleList(layers)


class _EncoderOnly(nn.Module):
    """Key names of MSDeformAttnTransformerEncoderOnly (msdeformattn.py:23-48): encoder.layers.{i}.*, level_embed."""

    def __init__(self, d_model, nhead, num_layers, d_ffn, n_levels, n_points=4):
        super().__init__()
        self.encoder = _Encoder([_EncoderLayer(d_model, d_ffn, n_levels, nhead, n_points) for _ in range(num_layers)])
        self.level_embed = nn.Parameter(torch.empty(n_levels, d_model))
        for n_, p in self.named_parameters():
            if p.dim() > 1 and "self_attn" not in n_:
                nn.init.xavier_uniform_(p)
        nn.init.normal_(self.level_embed)


class _ConvNorm(nn.Conv2d):
    """detectron2.layers.Conv2d key layout: weight[, bias], norm.{weight,bias}; forward = conv -> norm -> act."""

    def __init__(self, cin, cout, k, padding=0, bias=True, norm=None, act=None):
        super().__init__(cin, cout, k, padding=padding, bias=bias)
        self.norm = norm
        self.act = act

    def forward_cl(self, x_cl):
        """x_cl [N,H,W,Cin] channel-last -> [N,H,W,Cout] channel-last (conv -> norm -> act)."""
        N, H, W, Cin = x_cl.shape
        if self.kernel_size == (1, 1):
            y = nn_ops.linear(x_cl, self.weight.view(self.out_channels, Cin), self.bias)
        else:
            y = nn_ops.conv2d_cl(x_cl, self.weight, self.bias, padding=self.padding)
        if self.norm is not None:
            y = self.norm(y.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)        # GroupNorm on the NCHW view
        if self.act is not None:
            y = self.act(y)
        return y if y.is_contiguous() else y.contiguous()


@SEM_SEG_HEADS_REGISTRY.register()
class MSDeformAttnPixelDecoder(nn.Module):
    def __init__(self, input_shape, *args, **kwargs):
        super().__init__()
        if is_cfg(input_shape):                      # @configurable: (cfg, input_shape)
            kwargs = self.from_config(input_shape, *args)
            input_shape = kwargs.pop("input_shape")
        self._build(input_shape, **kwargs)

    @classmethod
    def from_config(cls, cfg, input_shape):
        """msdeformattn.py:296-314"""
        h = cfg.MODEL.SEM_SEG_HEAD
        return dict(input_shape={k: v for k, v in input_shape.items() if k in h.IN_FEATURES},
                    conv_dim=h.CONVS_DIM, mask_dim=h.MASK_DIM, norm=h.NORM,
                    transformer_dropout=cfg.MODEL.MASK_FORMER.DROPOUT, transformer_nheads=cfg.MODEL.MASK_FORMER.NHEADS,
                    transformer_dim_feedforward=1024, transformer_enc_layers=h.TRANSFORMER_ENC_LAYERS,
                    transformer_in_features=h.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES, common_stride=h.COMMON_STRIDE)

    def _build(self, input_shape, *, transformer_dropout, transformer_nheads, transformer_dim_feedforward,
               transformer_enc_layers, conv_dim, mask_dim, norm=None, transformer_in_features, common_stride):
        shapes = sorted(input_shape.items(), key=lambda kv: kv[1].stride)
        self.in_features = [k for k, _ in shapes]
        self.feature_channels = [v.channels for _, v in shapes]
        tshapes = [(k, v) for k, v in shapes if k in transformer_in_features]
        self.transformer_in_features = [k for k, _ in tshapes]
        tstrides = [v.stride for _, v in tshapes]
        self.transformer_num_feature_levels = len(tshapes)
        if conv_dim != 32 * transformer_nheads:
            raise ValueError("the B200 MSDeformAttn kernel requires head_dim == 32")
        self.input_proj = nn.ModuleList(
            nn.Sequential(nn.Conv2d(v.channels, conv_dim, kernel_size=1), nn.GroupNorm(32, conv_dim))
            for _, v in tshapes[::-1])
        for proj in self.input_proj:
            nn.init.xavier_uniform_(proj[0].weight, gain=1)
            nn.init.zeros_(proj[0].bias)
        self.transformer = _EncoderOnly(conv_dim, transformer_nheads, transformer_enc_layers,
                                        transformer_dim_feedforward, self.transformer_num_feature_levels)
        self.mask_dim, self.conv_dim = mask_dim, conv_dim
        self.mask_features = _ConvNorm(conv_dim, mask_dim, 1)
        self.maskformer_num_feature_levels = 3
        self.common_stride = common_stride
        self.num_fpn_levels = int(np.log2(min(tstrides)) - np.log2(common_stride))
        use_bias = norm == ""
        mk_norm = (lambda: nn.GroupNorm(32, conv_dim)) if norm == "GN" else (lambda: None)
        if norm not in ("", "GN", None):
            raise ValueError(f"unsupported norm {norm!r}")
        lateral, output = [], []
        for idx, cin in enumerate(self.feature_channels[:self.num_fpn_levels]):
            lat = _ConvNorm(cin, conv_dim, 1, bias=use_bias, norm=mk_norm())
            outc = _ConvNorm(conv_dim, conv_dim, 3, padding=1, bias=use_bias, norm=mk_norm(), act=F.relu)
            self.add_module(f"adapter_{idx + 1}", lat)
            self.add_module(f"layer_{idx + 1}", outc)
            lateral.append(lat)
            output.append(outc)
        for m in [self.mask_features] + lateral + output:
            nn.init.kaiming_uniform_(m.weight, a=1)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        self.lateral_convs, self.output_convs = lateral[::-1], output[::-1]
        self.eval()

    @torch.no_grad()
    def forward_features(self, features):
        """msdeformattn.py:316-360.  Frames are the batch dim; nothing here mixes frames."""
        tokens, poss, shapes = [], [], []
        for idx, f in enumerate(self.transformer_in_features[::-1]):      # res5, res4, res3
            x = features[f].float()
            n, cin, h, w = x.shape
            xt = x.permute(0, 2, 3, 1)                                      # channel-last view (copy only if NCHW)
            xt = (xt if xt.is_contiguous() else xt.contiguous()).view(n, h * w, cin)
            conv, gn = self.input_proj[idx][0], self.input_proj[idx][1]
            y = nn_ops.linear(xt, conv.weight.view(conv.out_channels, cin), conv.bias)     # 1x1 conv
            if nn_ops.fused_glue():                                          # channel-last GroupNorm, one pass
                y = nn_ops.groupnorm_cl(y.view(n, h, w, -1), gn)[0].view(n, h * w, -1)
            else:
                y = gn(y.transpose(1, 2)).transpose(1, 2)                    # GroupNorm(32) per frame
            shapes.append((h, w))
            tokens.append(y)                                                 # [N,hw,C]
            poss.append(position.sine_2d(h, w, y.device, y.shape[-1] // 2) + self.transformer.level_embed[idx])
        src = torch.cat(tokens, 1).contiguous()
        pos = torch.cat(poss, 0).unsqueeze(0)                               # [1,Len,C] (frame-invariant)
        starts = [0]
        for h, w in shapes[:-1]:
            starts.append(starts[-1] + h * w)
        layers = self.transformer.encoder.layers
        carry = None
        for i, layer in enumerate(layers):
            src, carry = layer(src, pos, shapes, starts, carry, emit_next=i + 1 < len(layers))
        n = src.shape[0]
        out_cl = []
        for i, (h, w) in enumerate(shapes):
            out_cl.append(src[:, starts[i]:starts[i] + h * w].reshape(n, h, w, -1))     # channel-last [N,h,w,C]
        fused = nn_ops.fused_glue() and all(c.norm is not None for c in self.lateral_convs + self.output_convs)
        last_operand = None
        for idx, f in enumerate(self.in_features[:self.num_fpn_levels][::-1]):
            x = features[f].float().permute(0, 2, 3, 1)
            x = x if x.is_contiguous() else x.contiguous()
            if fused:
                y, last_operand = self._fpn_level_fused(idx, x, out_cl[-1], idx == self.num_fpn_levels - 1)
                out_cl.append(y)
                continue
            cur = self.lateral_convs[idx].forward_cl(x)
            up = F.interpolate(out_cl[-1].permute(0, 3, 1, 2), size=cur.shape[1:3], mode="bilinear", align_corners=False)
            out_cl.append(self.output_convs[idx].forward_cl(cur + up.permute(0, 2, 3, 1)))
        out = [o.permute(0, 3, 1, 2) for o in out_cl]                       # NCHW views of channel-last storage
        multi_scale = out[:self.maskformer_num_feature_levels]
        if last_operand is not None:      # the last GroupNorm already emitted the operand of the 1x1 mask-feature conv
            mf = self.mask_features
            mask_features = nn_ops.linear_prepped(last_operand, mf.weight.view(mf.out_channels, -1), mf.bias)
            mask_features = mask_features.permute(0, 3, 1, 2)
        else:
            mask_features = self.mask_features.forward_cl(out_cl[-1]).permute(0, 3, 1, 2)
        return mask_features, out[-1], out[0], multi_scale

    def _fpn_level_fused(self, idx, x_cl, coarser_cl, emit_operand):
        """One top-down FPN level (msdeformattn.py:345-354) with the GroupNorm glue fused (csrc/groupnorm.cu):
        lateral 1x1 GEMM -> [stats] -> GN + bilinear(coarser) add, written straight as the zero-padded operand of the 3x3
        convolution -> nine tap GEMMs -> [stats] -> GN + ReLU -> fp32 level output (+ the operand of the mask-feature conv)."""
        lat, outc = self.lateral_convs[idx], self.output_convs[idx]
        N, H, W, Cin = x_cl.shape
        cur = nn_ops.linear(x_cl, lat.weight.view(lat.out_channels, Cin), lat.bias)          # [N,H,W,C] fp32
        if nn_ops.splitting():
            _, operand = nn_ops.groupnorm_cl(cur, lat.norm, lowres=coarser_cl, want_f32=False, for_gemm=True,
                                             pad=outc.padding[0])
            y = nn_ops.conv2d_cl_operand(operand, H, W, outc.weight, outc.bias)               # view of the padded rows
        else:
            summed, _ = nn_ops.groupnorm_cl(cur, lat.norm, lowres=coarser_cl)
            y = nn_ops.conv2d_cl(summed, outc.weight, outc.bias, padding=outc.padding)
            y = y if (y.stride(3) == 1 and y.stride(2) == y.shape[3]) else y.contiguous()
        return nn_ops.groupnorm_cl(y, outc.norm, relu=True, for_gemm=emit_operand)
