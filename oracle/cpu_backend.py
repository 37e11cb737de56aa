"""TEST / BASELINE INFRASTRUCTURE (tests/, bench.py cpu_baseline and --impl reference legs only): runs the product's host-side model with every hot-path operator replaced by its CPU oracle
(oracle/ops_ref.py), by monkeypatching `univs_b200.ops` for the duration of a `with` block.  This lets the host
logic (addressing, layouts, prompt plumbing, state_dict mapping) be checked against the reference on a machine
without a GPU.  The product package itself contains no such switch and no CPU code path."""
import contextlib

import torch

from oracle import ops_ref
from univs_b200 import ops


def unpack_bits(bits, n):
    b = bits.to(torch.int64) & 0xFFFFFFFF
    return ((b.unsqueeze(-1) >> torch.arange(32, device=bits.device)) & 1).flatten(-2)[..., :n].to(torch.uint8)


def _swin(qkv, qkv_bias, table, num_heads, window, shift, precision=None):
    return ops_ref.swin_window_attention(qkv, qkv_bias, table, num_heads, window, shift)


def _msda_enc(value, shapes, starts, offs_logits, num_levels=3, num_points=4, tile=None, value_bias=None,
              offs_logits_bias=None, split=None):
    # the reference adds the biases in the producing linears (ms_deform_attn.py:98-105)
    if value_bias is not None:
        value = value + value_bias.view(1, 1, value.shape[2], value.shape[3])
    if offs_logits_bias is not None:
        offs_logits = offs_logits + offs_logits_bias
    y = ops_ref.ms_deform_attn_fused(value, shapes, starts, offs_logits, value.shape[2], num_levels, num_points)
    return _maybe_split(y, split) if split else y


def _msda_fwd(value, shapes, starts, loc, w):
    sh = shapes.tolist() if torch.is_tensor(shapes) else shapes
    st = starts.tolist() if torch.is_tensor(starts) else starts
    return ops_ref.ms_deform_attn(value, sh, st, loc, w)


def _einsum(mask_embed, feats_cl, out=None, mode=None, tag=None):
    r = ops_ref.mask_einsum(mask_embed, feats_cl.transpose(1, 2))
    if out is not None:
        out.copy_(r)
        return out
    return r.contiguous()


def _bits(mask_logits, hw, target_hw):
    m = ops_ref.attn_mask_from_logits(mask_logits, hw, target_hw)          # [T,Q,S] uint8
    return ops.pack_mask_bits(m.bool()), (~m.bool().all(-1)).to(torch.int32)


def _bits_direct(mask_logits):
    m = ops_ref.attn_mask_direct(mask_logits)
    return ops.pack_mask_bits(m.bool()), (~m.bool().all(-1)).to(torch.int32)


def _mha(q, k, v, mask_bits=None, row_open=None, precision=None):
    mask = None
    if mask_bits is not None:
        mask = unpack_bits(mask_bits, k.shape[1])
    return ops_ref.mha_core(q, k, v, q.shape[-1] // 32, mask, unmask_full_rows=row_open is not None)


def _proca(q, ks, vs, km, vm):
    return ops_ref.proca_core(q, ks, vs, km, vm, q.shape[-1] // 32)


def _chunk(K):
    return K


def _split(x, chunk=None):
    b = x.contiguous().view(torch.int32)
    mag = ((b & 0x7FFFFFFF) + 0x1000) & -8192              # round-to-nearest (ties away), cvt.rna.tf32.f32
    hi = (mag | (b & -2147483648)).view(torch.float32)
    C = x.shape[-1]
    kc = _chunk(C) if chunk is None else chunk
    hl = torch.stack([hi.reshape(*x.shape[:-1], C // kc, kc), (x - hi).reshape(*x.shape[:-1], C // kc, kc)], -2)
    return hl.reshape(*x.shape[:-1], 2 * C)


def _split16(x, scaled):
    hi = x.clamp(-65504, 65504).half()
    lo = ((x - hi.float()) * (2048.0 if scaled else 1.0)).clamp(-65504, 65504).half()
    if not scaled:
        return torch.cat([hi, lo], -1)
    from univs_b200.ops import f16_chunk
    C = x.shape[-1]
    kc = f16_chunk(C)
    hs = (hi.float() / 2048.0).half()
    parts = [t.reshape(*x.shape[:-1], C // kc, kc) for t in (lo, hs, hi)]
    return torch.stack(parts, -2).reshape(*x.shape[:-1], 3 * C)


def _maybe_split(y, split):
    if not split:
        return y
    if split in ("f16", "f16u"):
        return _split16(y, split == "f16")
    if split == "f16c":                       # compact [hi | lo*2^11] (UNIVS_SPLIT_F16C)
        hi = y.clamp(-65504, 65504).half()
        lo = ((y - hi.float()) * 2048.0).clamp(-65504, 65504).half()
        return torch.cat([hi, lo], -1)
    return _split(y)


def _layernorm(x, weight, bias, eps=1e-5, residual=None, want_sum=False, split=None, residual_bias=None):
    s = x if residual is None else x + residual
    if residual_bias is not None:
        s = s + residual_bias
    y = torch.nn.functional.layer_norm(s, (x.shape[-1],), weight, bias, eps)
    return (s if want_sum else None), _maybe_split(y, split)


def _groupnorm_cl(x, weight, bias, groups, eps=1e-5, lowres=None, relu=False, want_f32=True, split=None, pad=0,
                  out_split=None):
    y = ops_ref.groupnorm_cl(x, weight, bias, groups, eps, lowres, relu)
    op = None
    if split:
        op = _maybe_split(y, split)
        if pad:
            op = torch.nn.functional.pad(op, (0, 0, pad, pad, pad, pad))
        if out_split is not None:
            out_split.copy_(op)
            op = out_split
    return (y if want_f32 else None), op


def _layernorm_multi(x, weight, bias, eps=1e-5, residual=None, residual_bias=None, want_f32=True, split=None, pos=None,
                     want_operand=True):
    _, y = _layernorm(x, weight, bias, eps, residual, False, None, residual_bias)
    op = _maybe_split(y, split) if (split and want_operand) else None
    op_pos = None
    if pos is not None:
        C = x.shape[-1]
        p = pos.reshape(-1, C)
        yp = (y.reshape(-1, p.shape[0], C) + p).reshape(y.shape)
        op_pos = _maybe_split(yp, split)
    return (y if want_f32 else None), op, op_pos


def _gemm_tc(x16, x_offs, w16, w_offs, k, alpha=1.0, bias=None, addend=None, out=None, want_f32=True, want_operand=False,
             act=0, tap_rows=None, rows=None):
    y, y16 = ops_ref.gemm_f16x3(x16, x_offs, w16, w_offs, k, alpha, bias, addend, act, tap_rows, rows)
    if out is not None:
        out.copy_(y)
        y = out
    return (y if (want_f32 or out is not None) else None), (y16 if want_operand else None)


_PATCH = {"layernorm": _layernorm,
          "gemm_f16x3_tc": _gemm_tc,
          "layernorm_multi": _layernorm_multi,
          "groupnorm_cl": _groupnorm_cl,
          "patchify_normalize": lambda f, m, s, padded, patch=4, split=None: _maybe_split(
              ops_ref.patchify_normalize(f, m, s, padded, patch), split),
          "layernorm_merge2x2": lambda x, w, b, eps=1e-5, split=None: _maybe_split(
              ops_ref.layernorm_merge2x2(x, w, b, eps), split),
          # `compact` asks for the [hi | lo'] container where the kernel in use can write it; the consumer tells the two
          # containers apart by their width, so the oracle keeps the K-chunk one
          "swin_window_attention_operand": lambda q, b, t, nh, ws, sh, compact=False: _split16(_swin(q, b, t, nh, ws, sh), True),
          "gelu": lambda x, split=None, bias=None: _maybe_split(torch.nn.functional.gelu(x if bias is None else x + bias), split),
          "relu": lambda x, split=None, bias=None: _maybe_split(torch.relu(x if bias is None else x + bias), split),
          "split_tf32": _split,
          "split_operand": lambda x, split="tf32": _maybe_split(x, split),
          "swin_window_attention": _swin, "ms_deform_attn_encoder": _msda_enc, "ms_deform_attn_forward": _msda_fwd,
          "mask_einsum": _einsum, "attn_mask_bits": _bits, "attn_mask_bits_direct": _bits_direct,
          "mask_feature_pool": lambda f, hw, thw, mode=None: ops_ref.mask_feature_pool(f, hw, thw), "mha_core": _mha, "proca_core": _proca,
          "round_tf32": lambda x, out=None: x, "prepare_mask_features": lambda x, mode=None: x}


@contextlib.contextmanager
def oracle_ops(policy="fp32"):
    """`policy`: the nn_ops operand policy while the oracle is active.  "fp32" (default) = plain fp32 operands and
    IEEE GEMMs, i.e. the reference arithmetic; "tf32x3" exercises the split-operand plumbing with exact CPU GEMMs
    (the fp16 GEMM path has no CPU implementation)."""
    from univs_b200 import nn_ops
    saved = {k: getattr(ops, k) for k in _PATCH}
    saved_policy = nn_ops.policy()
    try:
        for k, f in _PATCH.items():
            setattr(ops, k, f)
        nn_ops.set_policy(policy)
        yield
    finally:
        for k, f in saved.items():
            setattr(ops, k, f)
        nn_ops.set_policy(saved_policy)
