"""TEST INFRASTRUCTURE ONLY -- CPU (torch fp32/fp64) restatement of the arithmetic of
every hot-path operator, at exactly the operator boundary the C-ABI in
include/univs_b200.h exposes.  Each function cites the reference file:line it
restates (paths relative to the reference tree).

Parity pin: every function here is checked against the reference's OWN source files
(loaded by path through oracle/ref_shim.py) in tests/test_oracle_vs_reference.py, and
against the committed golden vectors in tests/golden/ (generated from the reference by
tests/golden/make_golden.py).  The reference ships no golden vectors of its own for
this path except the MSDeformAttn shapes of ops/test.py:24-63, which
tests/test_oracle_vs_reference.py::test_msda_reference_test_shapes replays.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------
# a7: MSDeformAttn core.  ops/src/cuda/ms_deform_im2col_cuda.cuh:38-89 (bilinear),
# :242-304 (per-output accumulation), h_im = loc_h*H - 0.5 (:290-291), zero padding
# outside (-1, H) x (-1, W) (:293).
# ---------------------------------------------------------------------------
def ms_deform_attn(value, spatial_shapes, level_start_index, sampling_loc, attn_weight):
    """value [N,S,M,D]; spatial_shapes [[H,W]..]; sampling_loc [N,Lq,M,L,P,2] (x,y in [0,1]);
    attn_weight [N,Lq,M,L,P]  ->  [N,Lq,M*D]"""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_loc.shape
    shapes = [(int(h), int(w)) for h, w in spatial_shapes]
    starts = [int(s) for s in level_start_index]
    out = value.new_zeros(N, Lq, M, D)
    n_idx = torch.arange(N, device=value.device).view(N, 1, 1, 1)
    m_idx = torch.arange(M, device=value.device).view(1, 1, M, 1)
    for l, (H, W) in enumerate(shapes):
        v = value[:, starts[l]:starts[l] + H * W]                  # [N,HW,M,D]
        x = sampling_loc[:, :, :, l, :, 0] * W - 0.5                # [N,Lq,M,P]
        y = sampling_loc[:, :, :, l, :, 1] * H - 0.5
        inside = (y > -1) & (x > -1) & (y < H) & (x < W)
        x0 = torch.floor(x); y0 = torch.floor(y)
        lx = x - x0; ly = y - y0
        acc = value.new_zeros(N, Lq, M, P, D)
        for dy, wy in ((0, 1 - ly), (1, ly)):
            for dx, wx in ((0, 1 - lx), (1, lx)):
                xi = (x0 + dx).long(); yi = (y0 + dy).long()
                ok = inside & (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
                lin = (yi.clamp(0, H - 1) * W + xi.clamp(0, W - 1))  # [N,Lq,M,P]
                g = v[n_idx, lin, m_idx]                              # [N,Lq,M,P,D]
                acc = acc + g * (wy * wx * ok).unsqueeze(-1)
        out = out + (acc * attn_weight[:, :, :, l].unsqueeze(-1)).sum(3)
    return out.reshape(N, Lq, M * D)


def encoder_reference_points(spatial_shapes):
    """msdeformattn.py:143-155 with valid_ratios == 1 (masks are all-False, :62):
    pixel centres normalised by the level size; identical for every level -> [Len, 2] (x,y)."""
    pts = []
    for H, W in spatial_shapes:
        H = int(H); W = int(W)
        ys = (torch.arange(H, dtype=torch.float32) + 0.5) / H
        xs = (torch.arange(W, dtype=torch.float32) + 0.5) / W
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        pts.append(torch.stack([xx.reshape(-1), yy.reshape(-1)], -1))
    return torch.cat(pts, 0)


def ms_deform_attn_fused(value, spatial_shapes, level_start_index, offs_logits, M, L, P):
    """Fused front end of MSDeformAttn.forward (ops/modules/ms_deform_attn.py:98-117) for the
    encoder case (queries == pixels of the level pyramid, reference points = pixel centres):
      offs_logits [N,Len,M*L*P*3] = concat(sampling_offsets(q) [M,L,P,2], attention_weights(q) [M,L*P])
      softmax over L*P (:101-102); loc = ref + off / (W_l, H_l) (:105-108)."""
    N, Len, _ = offs_logits.shape
    n_off = M * L * P * 2
    off = offs_logits[..., :n_off].reshape(N, Len, M, L, P, 2)
    logit = offs_logits[..., n_off:].reshape(N, Len, M, L * P)
    w = torch.softmax(logit, -1).reshape(N, Len, M, L, P)
    ref = encoder_reference_points(spatial_shapes).to(value)            # [Len,2]
    norm = torch.tensor([[int(w_), int(h_)] for h_, w_ in spatial_shapes], dtype=value.dtype, device=value.device)
    loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    return ms_deform_attn(value, spatial_shapes, level_start_index, loc, w)


# ---------------------------------------------------------------------------
# a2/a3: Swin (shifted-)window attention with the partition/roll/pad addressing folded in.
# swin.py:131-171 (attention), :247-289 (pad -> roll -> partition ... reverse -> roll -> crop),
# :413-440 (shift mask on the padded grid, -100 not -inf).
# ---------------------------------------------------------------------------
def swin_window_attention(qkv, qkv_bias, rel_bias_table, num_heads, window, shift, return_scores=False):
    """qkv [B,H,W,3C] = LN(x) @ Wqkv^T on the UNPADDED token grid WITHOUT the bias (the operator adds
    qkv_bias to every token); pad tokens are zeros after norm1 in the reference (swin.py:247-255),
    so their qkv equals the qkv bias.
    rel_bias_table [(2w-1)^2, nH].  Returns the attention output (pre-proj) [B,H,W,C]
    (with return_scores also the biased, masked pre-softmax scores [B, nW, nH, N, N])."""
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    ws = window
    d = C // num_heads
    Hp = (H + ws - 1) // ws * ws
    Wp = (W + ws - 1) // ws * ws
    full = qkv_bias.view(1, 1, 1, C3).expand(B, Hp, Wp, C3).clone()
    full[:, :H, :W] = qkv + qkv_bias
    if shift > 0:
        full = torch.roll(full, shifts=(-shift, -shift), dims=(1, 2))
    nWh, nWw = Hp // ws, Wp // ws
    win = full.view(B, nWh, ws, nWw, ws, 3, num_heads, d).permute(5, 0, 1, 3, 6, 2, 4, 7)
    win = win.reshape(3, B, nWh * nWw, num_heads, ws * ws, d)
    q, k, v = win[0] * (d ** -0.5), win[1], win[2]
    attn = q @ k.transpose(-2, -1)                                    # [B,nW,nH,N,N]
    # relative position bias (swin.py:108-121, 148-156)
    ar = torch.arange(ws, device=qkv.device)
    cy, cx = torch.meshgrid(ar, ar, indexing="ij")
    cy = cy.reshape(-1); cx = cx.reshape(-1)
    idx = (cy[:, None] - cy[None, :] + ws - 1) * (2 * ws - 1) + (cx[:, None] - cx[None, :] + ws - 1)
    bias = rel_bias_table[idx.reshape(-1)].view(ws * ws, ws * ws, num_heads).permute(2, 0, 1)
    attn = attn + bias[None, None]
    if shift > 0:
        lab = torch.zeros(Hp, Wp, device=qkv.device)
        cnt = 0
        for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
                lab[hs, wsl] = cnt
                cnt += 1
        lab = lab.view(nWh, ws, nWw, ws).permute(0, 2, 1, 3).reshape(nWh * nWw, ws * ws)
        m = (lab[:, None, :] != lab[:, :, None]).to(attn.dtype) * -100.0   # [nW,N,N]
        attn = attn + m[None, :, None]
    scores = attn
    attn = torch.softmax(attn, -1)
    o = attn @ v                                                      # [B,nW,nH,N,d]
    o = o.view(B, nWh, nWw, num_heads, ws, ws, d).permute(0, 1, 4, 2, 5, 3, 6).reshape(B, Hp, Wp, C)
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    o = o[:, :H, :W].contiguous()
    return (o, scores) if return_scores else o


# ---------------------------------------------------------------------------
# a11: mask einsum "btqc,btchw->btqhw" then transpose(1,2)
# (video_mask2former_transformer_decoder_univs.py:527-528)
# ---------------------------------------------------------------------------
def mask_einsum(mask_embed, mask_features):
    """mask_embed [T,Q,C], mask_features [T,C,HW] -> [Q,T,HW]"""
    return torch.einsum("tqc,tcp->qtp", mask_embed, mask_features)


def attn_mask_from_logits(mask_logits, hw, target_hw):
    """..._univs.py:555-566: bilinear resize (align_corners=False) of the [Q,T,H,W] logits to the
    next level's size, sigmoid < 0.5 -> True (= key blocked).  Returns uint8 [T,Q,h*w]
    (shared by the 8 heads -- the reference repeats it per head)."""
    Q, T, _ = mask_logits.shape
    H, W = hw
    m = F.interpolate(mask_logits.view(Q, T, H, W), size=target_hw, mode="bilinear", align_corners=False)
    m = m.permute(1, 0, 2, 3).reshape(T, Q, -1)
    return (m.sigmoid() < 0.5).to(torch.uint8)


def mask_feature_pool(feats_cl, hw, target_hw):
    """Pooled-feature variant (csrc/decoder_glue.cu): the bilinear resize of ..._univs.py:555-560 applied to the mask
    features [T, H*W, C] instead of the logits (resize and einsum are both linear).  Returns [T, h*w, C]."""
    T, _, C = feats_cl.shape
    H, W = hw
    x = feats_cl.view(T, H, W, C).permute(0, 3, 1, 2)
    y = F.interpolate(x, size=target_hw, mode="bilinear", align_corners=False)
    return y.permute(0, 2, 3, 1).reshape(T, -1, C).contiguous()


def attn_mask_direct(mask_logits):
    """[Q,T,S] logits already at the memory resolution -> uint8 [T,Q,S], 1 = blocked (sigmoid < 0.5, ..._univs.py:561)"""
    return (mask_logits.permute(1, 0, 2).sigmoid() < 0.5).to(torch.uint8)


# ---------------------------------------------------------------------------
# a12/a13: multi-head attention core of nn.MultiheadAttention (post in-projection,
# pre out-projection): torch.nn.functional.multi_head_attention_forward, called from
# transformer_layers.py:34-44 (self) and :95-115 (cross).  Boolean mask: True = blocked
# (-inf before softmax).  Row rule of ..._univs.py:390: a query row that is blocked
# everywhere is un-blocked everywhere.
# ---------------------------------------------------------------------------
def mha_core(q, k, v, num_heads, mask=None, unmask_full_rows=False):
    """q [B,Lq,C], k,v [B,Lk,C] (already projected; q NOT yet scaled), mask uint8/bool
    [B or 1, Lq, Lk] with 1 = blocked.  Returns [B,Lq,C]."""
    B, Lq, C = q.shape
    Lk = k.shape[1]
    d = C // num_heads
    qh = q.view(B, Lq, num_heads, d).transpose(1, 2) * (d ** -0.5)
    kh = k.view(B, Lk, num_heads, d).transpose(1, 2)
    vh = v.view(B, Lk, num_heads, d).transpose(1, 2)
    s = qh @ kh.transpose(-2, -1)                                     # [B,h,Lq,Lk]
    if mask is not None:
        mb = mask.bool()
        if unmask_full_rows:
            mb = mb & ~mb.all(-1, keepdim=True)
        s = s.masked_fill(mb[:, None], float("-inf"))
    p = torch.softmax(s, -1)
    return (p @ vh).transpose(1, 2).reshape(B, Lq, C)


# ---------------------------------------------------------------------------
# a14: ProCA attention core (..._univs.py:456-496): every prompt query (per frame)
# attends to [its own token ; its L prompt-memory tokens]; q-len 1.
# ---------------------------------------------------------------------------
def proca_core(q, k_self, v_self, k_mem, v_mem, num_heads):
    """q,k_self,v_self [P,T,C]; k_mem,v_mem [P,Tm,L,C] with Tm in {1,T}.  Returns [P,T,C]."""
    P, T, C = q.shape
    L = k_mem.shape[2]
    km = k_mem.expand(P, T, L, C)
    vm = v_mem.expand(P, T, L, C)
    k = torch.cat([k_self.unsqueeze(2), km], 2).reshape(P * T, 1 + L, C)
    v = torch.cat([v_self.unsqueeze(2), vm], 2).reshape(P * T, 1 + L, C)
    return mha_core(q.reshape(P * T, 1, C), k, v, num_heads).view(P, T, C)


# ---------------------------------------------------------------------------
# a10: sine position encodings
# ---------------------------------------------------------------------------
def pos2d_sine(h, w, num_pos_feats=128, temperature=10000.0):
    """mask2former/modeling/transformer_decoder/position_encoding.py:29-52, normalize=True -> [2*npf,h,w]"""
    eps, scale = 1e-6, 2 * math.pi
    y = torch.arange(1, h + 1, dtype=torch.float32).view(h, 1).expand(h, w)
    x = torch.arange(1, w + 1, dtype=torch.float32).view(1, w).expand(h, w)
    y = y / (h + eps) * scale
    x = x / (w + eps) * scale
    dim_t = torch.arange(num_pos_feats, dtype=torch.float32)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_feats)
    px = x[:, :, None] / dim_t
    py = y[:, :, None] / dim_t
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), 3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), 3).flatten(2)
    return torch.cat((py, px), 2).permute(2, 0, 1)


def pos3d_sine_arbitrary_t(frame_indices, h, w, num_pos_feats=128, temperature=10000.0, num_max_frames=128):
    """univs/modeling/transformer_decoder/position_encoding.py:142-169 -> [T, 2*npf, h, w]"""
    base = pos2d_sine(h, w, num_pos_feats, temperature)              # [C,h,w] (same y/x terms)
    z = frame_indices.to(torch.float32) / num_max_frames * (2 * math.pi)   # [T]
    dim_z = torch.arange(num_pos_feats * 2, dtype=torch.float32)
    dim_z = temperature ** (2 * torch.div(dim_z, 2, rounding_mode="trunc") / (num_pos_feats * 2))
    pz = z[:, None] / dim_z
    pz = torch.stack((pz[:, 0::2].sin(), pz[:, 1::2].cos()), 2).flatten(1)   # [T,C]
    return base[None] + pz[:, :, None, None]


# ---------------------------------------------------------------------------
# fused row-wise kernels around the GEMMs (csrc/elementwise.cu)
# ---------------------------------------------------------------------------
def _chunk(K):
    return K


def split_tf32(x, chunk=None):
    """[..., C] -> [..., 2C] in K-chunks [hi_c | lo_c]: hi = x rounded to nearest TF32 (cvt.rna.tf32.f32), lo = x - hi."""
    b = x.contiguous().view(torch.int32)
    mag = ((b & 0x7FFFFFFF) + 0x1000) & -8192              # round-to-nearest (ties away), cvt.rna.tf32.f32
    hi = (mag | (b & -2147483648)).view(torch.float32)
    C = x.shape[-1]
    kc = _chunk(C) if chunk is None else chunk
    hl = torch.stack([hi.reshape(*x.shape[:-1], C // kc, kc), (x - hi).reshape(*x.shape[:-1], C // kc, kc)], -2)
    return hl.reshape(*x.shape[:-1], 2 * C)


def layernorm(x, weight, bias, eps=1e-5, residual=None, residual_bias=None):
    """nn.LayerNorm over the last dim of (x + residual + residual_bias) (swin.py:246,292; transformer_layers.py:42)."""
    s = x if residual is None else x + residual
    if residual_bias is not None:
        s = s + residual_bias
    return s, F.layer_norm(s.double(), (x.shape[-1],), weight.double(), bias.double(), eps).float()


def gelu(x, bias=None):
    """nn.GELU() default = exact erf form (swin.py:24-41), applied to x + bias."""
    return F.gelu((x if bias is None else x + bias).double()).float()


# --------------------------------------------------------------------------------------------------------------
# Fused glue around the library GEMMs (csrc/groupnorm.cu, csrc/swin_glue.cu): the reference's own op sequences
# --------------------------------------------------------------------------------------------------------------
def groupnorm_cl(x_cl, weight, bias, groups, eps=1e-5, lowres_cl=None, relu=False):
    """detectron2 Conv2d tail on channel-last data: GroupNorm (msdeformattn.py:218-221, :249-283) [+ the top-down add
    of the bilinearly upsampled coarser level, :345-354: F.interpolate(..., mode="bilinear", align_corners=False)]
    [+ ReLU].  x_cl [N,H,W,C], lowres_cl [N,h2,w2,C] -> [N,H,W,C]."""
    x = x_cl.permute(0, 3, 1, 2)
    y = torch.nn.functional.group_norm(x, groups, weight, bias, eps)
    if lowres_cl is not None:
        y = y + torch.nn.functional.interpolate(lowres_cl.permute(0, 3, 1, 2), size=x.shape[-2:], mode="bilinear",
                                                align_corners=False)
    if relu:
        y = torch.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def patchify_normalize(frames, pixel_mean, pixel_std, padded_size, patch=4):
    """(x - mean) / std (univs_prompt.py:165-168), zero pad right / bottom (ImageList.from_tensors), then the im2col of
    the stride-`patch` PatchEmbed convolution (swin.py:456-495): [N,3,H,W] -> [N, Hp/p, Wp/p, 3*p*p], column order
    (c, ky, kx) = proj.weight.view(E, -1)."""
    mean = torch.as_tensor(pixel_mean, dtype=torch.float32, device=frames.device).view(1, 3, 1, 1)
    std = torch.as_tensor(pixel_std, dtype=torch.float32, device=frames.device).view(1, 3, 1, 1)
    x = (frames.float() - mean) / std
    N, _, H, W = x.shape
    Hp, Wp = padded_size
    x = torch.nn.functional.pad(x, (0, Wp - W, 0, Hp - H))
    cols = torch.nn.functional.unfold(x, kernel_size=patch, stride=patch)          # [N, 3*p*p, L]
    return cols.transpose(1, 2).reshape(N, Hp // patch, Wp // patch, 3 * patch * patch).contiguous()


def layernorm_merge2x2(x_cl, weight, bias, eps=1e-5):
    """PatchMerging up to the reduction linear (swin.py:312-335): pad odd sizes, concatenate the 2x2 neighbours in the
    order x0, x1, x2, x3 = (0,0), (1,0), (0,1), (1,1), LayerNorm(4C)."""
    N, H, W, C = x_cl.shape
    x = torch.nn.functional.pad(x_cl, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
    return torch.nn.functional.layer_norm(x, (4 * C,), weight, bias, eps)


def gemm_f16x3(x16, x_offs, w16, w_offs, k, alpha=1.0, bias=None, addend=None, act=0, tap_rows=None, rows=None):
    """CPU restatement of univs_gemm_f16x3_tc (csrc/gemm_tc.cu): the dense layer the reference computes with nn.Linear + the
    activation / residual behind it (swin.py:35-41 Mlp, :138-169 qkv / proj; ms_deform_attn.py:98-120;
    transformer_layers.py).  Operands are fp16 pairs, value = hi + lo' * 2^-11.  Returns (y fp32, y as the compact operand
    [hi | lo' ] fp16)."""
    def value(t, offs, width=None, col0=0):
        width = k if width is None else width
        return (t[:, offs[0] + col0: offs[0] + col0 + width].double()
                + t[:, offs[1] + col0: offs[1] + col0 + width].double() * 2.0 ** -11)
    if tap_rows is None:
        y = (value(x16, x_offs) @ value(w16, w_offs).t() * alpha).float()
    else:       # univs_gemm_f16x3_tc_taps: y[m] = sum_t x[m + tap_rows[t]] w_t^T, rows beyond the end of x16 read as zeros
        xv = value(x16, x_offs)
        xv = torch.cat([xv, xv.new_zeros(int(rows) + max(tap_rows) - xv.shape[0], k)]) if int(rows) + max(tap_rows) > xv.shape[0] else xv
        acc = xv.new_zeros(int(rows), w16.shape[0])
        for t, r0 in enumerate(tap_rows):
            acc += xv[r0: r0 + int(rows)] @ value(w16, w_offs, k, t * k).t()
        y = (acc * alpha).float()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = torch.nn.functional.gelu(y)
    elif act == 2:
        y = torch.relu(y)
    if addend is not None:
        y = y + addend
    hi = y.clamp(-65504.0, 65504.0).half()
    lo = ((y - hi.float()) * 2048.0).clamp(-65504.0, 65504.0).half()
    return y, torch.cat([hi, lo], 1)
