"""Dense-layer plumbing around the library GEMMs, parameterised by the arithmetic policy (precision.py).

Policy "tf32x3" (strict, default): every GEMM / convolution runs on the TF32 tensor cores (cuBLAS / cuDNN through
torch) but on *split* operands, X*W^T = [Xh|Xl]*[Wh|Wh]^T + Xh*Wl^T, where hi = upper 19 bits (exact in TF32, so the
tensor core's truncation loses nothing) and lo = x - hi: products are fp32-equivalent (the dropped Xl*Wl term is
2^-22 relative), accumulation is fp32.  Activations are produced already split by the fused kernels that feed the
GEMMs (LayerNorm, GELU; csrc/elementwise.cu); weights are split once and cached.
Policy "tf32": plain operands, single TF32 GEMM.  Policy "fp32": plain operands, IEEE fp32 GEMMs (slow reference)."""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from . import ops

_policy = "fp32"
_wcache = {}


def set_policy(p: str):
    global _policy
    assert p in ("fp32", "tf32", "tf32x3")
    _policy = p
    _wcache.clear()


def policy() -> str:
    return _policy


def splitting() -> bool:
    return _policy == "tf32x3"


def _split_weight(w: torch.Tensor):
    """-> (W2 = [Wh | Wh] [N,2K], Wl [N,K]) for a 2-D weight, cached per parameter version."""
    key = id(w)
    ent = _wcache.get(key)
    if ent is None or ent[0] != (w.data_ptr(), w._version, tuple(w.shape)):
        w2d = w.detach().reshape(w.shape[0], -1).contiguous()
        hl = ops.split_tf32(w2d)                          # [N, 2K] = [hi | lo]
        K = w2d.shape[1]
        hi, lo = hl[:, :K], hl[:, K:].contiguous()
        ent = ((w.data_ptr(), w._version, tuple(w.shape)), torch.cat([hi, hi], 1).contiguous(), lo)
        _wcache[key] = ent
    return ent[1], ent[2]


def prep(x):
    """activation -> GEMM operand format of the active policy"""
    return ops.split_tf32(x.contiguous()) if splitting() else x


def layernorm(x, norm, residual=None, want_sum=False, for_gemm=True):
    """(sum | None, LN(x + residual)); the LN output is split iff it feeds a GEMM under the tf32x3 policy."""
    return ops.layernorm(x.contiguous(), norm.weight, norm.bias, norm.eps,
                         None if residual is None else residual.contiguous(), want_sum, for_gemm and splitting())


def gelu(x, for_gemm=True):
    return ops.gelu(x.contiguous(), for_gemm and splitting())


def relu(x, for_gemm=True):
    return ops.relu(x.contiguous(), for_gemm and splitting())


def linear_prepped(h, weight, bias=None, cache=True):
    """h: operand from prep()/layernorm()/gelu() ([..., 2K] when splitting, else [..., K])."""
    if not splitting():
        return F.linear(h, weight, bias)
    K = weight.shape[1]
    if cache:
        w2, wlo = _split_weight(weight)
    else:
        hl = ops.split_tf32(weight.contiguous())
        w2, wlo = torch.cat([hl[:, :K], hl[:, :K]], 1), hl[:, K:]
    y = F.linear(h, w2, bias)                                        # Xh*Wh + Xl*Wh (+ bias)
    y.view(-1, y.shape[-1]).addmm_(h.reshape(-1, 2 * K)[:, :K], wlo.t())   # + Xh*Wl
    return y


def linear(x, weight, bias=None, cache=True):
    return linear_prepped(prep(x), weight, bias, cache)


def conv2d_cl(x_cl, weight, bias=None, padding=0):
    """Convolution on a channel-last activation [N,H,W,Cin] -> [N,H,W,Cout] (storage channel-last)."""
    N, H, W, Cin = x_cl.shape
    if not splitting():
        y = F.conv2d(x_cl.permute(0, 3, 1, 2), weight, bias, padding=padding)
        return y.permute(0, 2, 3, 1)
    key = ("conv", id(weight))
    ent = _wcache.get(key)
    if ent is None or ent[0] != (weight.data_ptr(), weight._version):
        w = weight.detach()
        hl = ops.split_tf32(w.permute(0, 2, 3, 1).contiguous())            # [Cout,kh,kw,2Cin]
        hi, lo = hl[..., :Cin], hl[..., Cin:]
        w2 = torch.cat([hi, hi], -1).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        wl = lo.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        ent = ((weight.data_ptr(), weight._version), w2, wl)
        _wcache[key] = ent
    xs = ops.split_tf32(x_cl.contiguous())                                   # [N,H,W,2Cin]
    xs_nchw = xs.permute(0, 3, 1, 2)
    y = F.conv2d(xs_nchw, ent[1], bias, padding=padding)
    y = y + F.conv2d(xs_nchw[:, :Cin], ent[2], None, padding=padding)
    return y.permute(0, 2, 3, 1)


@contextlib.contextmanager
def ieee_fp32():
    """Small matmuls / convs that stay on plain torch ops (patch embedding, prompt pooling) under exact fp32."""
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    if _policy == "tf32":
        yield
        return
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32 = a
        torch.backends.cudnn.allow_tf32 = b
