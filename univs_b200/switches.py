"""Opt-in switches of paths that are built but not yet validated on a B200.

One place to read them: environment variable `UNIVS_<NAME>` wins, then `univs_b200/tuned.json` (committed once a path
has been validated and measured on the GPU -- flipping a default is a one-line edit there), then the built-in default,
which is always the path measured in round 1.  `active()` lists what is on, for the bench line."""
from __future__ import annotations

import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULTS = {
    "FUSED_GLUE": 0,        # nn_ops: channel-last GroupNorm + FPN glue, PatchMerging gather-LN, fused frame ingest
    "MSDA_TILE": 0,         # ops: tiled MSDeformAttn encoder kernel (tile width, 0 = untiled)
    "WIN_TC": 0,            # ops: tcgen05 window attention for 12x12 windows
    "MHA_TC": 0,            # ops: tcgen05 cross-attention (1; 3 = transposed-V diagnostic variant)
    "ROWWISE_V2": 0,        # csrc/elementwise.cu: bit 0 = 8-wide GELU / ReLU / operand split, bit 1 = wide-store LayerNorm
    "GEMM_TC": 0,           # nn_ops: dense layers on the own tcgen05 GEMM with fused epilogues (csrc/gemm_tc.cu)
    "CONV_FUSED": 0,        # nn_ops: k x k convolutions as one shifted-row accumulation of the own GEMM (taps) instead of k*k GEMMs
    "EINSUM_MC": 0,         # ops: cluster / TMA-multicast mask einsum (E resident per CTA pair)
    "MLP_CHUNK_MB": 0,      # backbone: Swin MLP in row chunks whose fp32 hidden + operand stay in L2 (0 = whole tensor)
    "POOLED_MASKS": 0,      # decoder: intermediate heads from pooled mask features
    "SHARD_DECODER": 0,     # meta_arch: frame-sharded decoder with token exchange (N > 1)
    "FRAME_STREAMS": 1,     # meta_arch: frame groups on CUDA streams (N == 1)
}

try:
    with open(os.path.join(_HERE, "tuned.json")) as f:
        _TUNED = {k.upper(): int(v) for k, v in json.load(f).items()}
except FileNotFoundError:
    _TUNED = {}


def get(name: str) -> int:
    name = name.upper()
    env = os.environ.get("UNIVS_" + name)
    if env is not None and env != "":
        return int(env)
    return int(_TUNED.get(name, DEFAULTS[name]))


def active() -> dict:
    """switches whose value differs from the round-1 default"""
    return {k: get(k) for k in DEFAULTS if get(k) != DEFAULTS[k]}


def export_native():
    """the C side reads UNIVS_ROWWISE_V2 with getenv at its first launch: make tuned.json visible to it"""
    if get("ROWWISE_V2") and os.environ.get("UNIVS_ROWWISE_V2") is None:
        os.environ["UNIVS_ROWWISE_V2"] = str(get("ROWWISE_V2"))
