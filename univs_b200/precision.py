"""Arithmetic policy of the path.

"tf32x3" (default, strict): fp32-equivalent everywhere.  Library GEMMs / convs run on the TF32 tensor cores over
         hi|lo-split operands (nn_ops.py), the attention contractions and the mask einsum use the 3xTF32 split
         in-kernel; fp32 accumulation, softmax, LayerNorm, residuals.  Meets the 1e-3 mask-logit bound
         (profiles/parity_at_scale_*.json).
"fp16x3": same construction with fp16 hi|lo*2^11 operands (fp16 has TF32's 11-bit significand, so hi*hi is exact in
         fp32 and two terms give 22 bits) at twice the tensor-core rate and half the operand bytes; operands
         saturate at +-65504 (LayerNorm / attention / activation outputs are far inside that range).
"tf32":  single-pass TF32 everywhere (round-to-nearest operands); ~1e-3 feature error which the discontinuous
         masked-attention decoder amplifies -- fails the mask-logit bound on random-init models; kept as the
         throughput upper bound of the same kernels.
"fp32":  plain operands, IEEE fp32 library GEMMs (CUDA-core SGEMM), 3xTF32 in-kernel; the slow reference policy.
SURVEY.md 7.3 measured that BF16 operands fail the bound outright, so no bf16 policy exists."""
from __future__ import annotations

import torch

from . import nn_ops, ops

_mode = "fp32"


def set_precision(mode: str):
    global _mode
    if mode == "fp32":
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ops.set_attention_precision(ops.PREC_TF32X3)
        ops.set_einsum_mode("mma3x")
    elif mode == "tf32":
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        ops.set_attention_precision(ops.PREC_TF32)
        ops.set_einsum_mode("tf32")
    elif mode in ("tf32x3", "fp16x3"):
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        ops.set_attention_precision(ops.PREC_TF32X3)
        ops.set_einsum_mode("f16x3" if mode == "fp16x3" else "mma3x")
    else:
        raise ValueError(f"unknown precision mode {mode!r} (fp16x3 | tf32x3 | tf32 | fp32)")
    nn_ops.set_policy(mode)
    _mode = mode


def get_precision() -> str:
    return _mode


# The documented default ("fp32": IEEE fp32 library GEMMs / convolutions, 3xTF32 in-kernel) must hold for a caller that never
# calls set_precision(): torch ships cudnn.allow_tf32 = True, which would run the plain-torch convolutions of that policy in
# 1-pass TF32 -- the one arithmetic that fails the 1e-3 mask-logit bound.
set_precision("fp32")
