"""Arithmetic policy of the path.

"fp32": library GEMMs/convs in full fp32 (cuBLAS/cuDNN TF32 disabled), attention contractions 3xTF32 (fp32-equivalent
        products, fp32 accumulate).  Strict-parity mode.
"tf32": library GEMMs/convs on TF32 tensor cores, attention contractions single TF32 (round-to-nearest operands),
        fp32 accumulate, fp32 softmax / LayerNorm / residuals.
The mask einsum always rounds its operands to nearest TF32 and accumulates in fp32 (see csrc/mask_einsum*.cu)."""
from __future__ import annotations

import torch

from . import ops

_mode = "fp32"


def set_precision(mode: str):
    global _mode
    if mode == "fp32":
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ops.set_attention_precision(ops.PREC_TF32X3)
    elif mode == "tf32":
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
        ops.set_attention_precision(ops.PREC_TF32)
    else:
        raise ValueError(f"unknown precision mode {mode!r} (fp32 | tf32)")
    _mode = mode


def get_precision() -> str:
    return _mode
