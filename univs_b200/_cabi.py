"""ctypes binding of the C ABI declared in include/univs_b200.h (one symbol per kernel entry point)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libunivs_b200.so")

_vp, _i, _i64 = C.c_void_p, C.c_int, C.c_int64

# name -> (restype, argtypes); must list every function declared in include/univs_b200.h
SIGNATURES = {
    "univs_b200_last_error": (C.c_char_p, []),
    "univs_b200_abi_version": (_i, []),
    "univs_ms_deform_attn_forward_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "univs_ms_deform_attn_backward_f32": (_i, []),
    "univs_ms_deform_attn_encoder_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "univs_ms_deform_attn_encoder_tiled_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "univs_swin_window_attention_f32": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "univs_swin_window_attention_f16x3out": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "univs_swin_window_attention_tc": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "univs_mask_einsum_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "univs_mask_einsum_f16x3": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "univs_mask_einsum_f16x3_cluster": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "univs_mask_einsum_mma_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "univs_attn_mask_bits_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "univs_mask_feature_pool_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i]),
    "univs_attn_mask_bits_direct_f32": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "univs_mha_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "univs_mha_forward_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "univs_mha_tc_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "univs_mha_tc_forward_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "univs_gemm_f16x3_tc": (_i, [_vp, _vp, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _i, _i, C.c_float, _vp, _vp, _i64,
                                 _vp, _i64, _vp, _i64, _i64, _i]),
    "univs_gemm_f16x3_tc_taps": (_i, [_vp, _vp, _i64, _i64, _i64, _i64, _vp, _i64, _i64, _i64, _i64, _i, _i, _i, _vp, C.c_float,
                                      _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i64, _i]),
    "univs_proca_forward_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "univs_layernorm_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, C.c_float, _vp, _vp, _i]),
    "univs_gelu_f32": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i]),
    "univs_relu_f32": (_i, [_vp, _vp, _vp, _i64, _i, _vp, _i]),
    "univs_split_tf32_f32": (_i, [_vp, _vp, _i64, _i, _i, _vp]),
    "univs_round_tf32_f32": (_i, [_vp, _vp, _vp, _i64]),
    "univs_groupnorm_workspace_bytes": (_i64, [_i, _i, _i, _i]),
    "univs_groupnorm_stats_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i64, _i64, _i, C.c_float, _vp, _vp]),
    "univs_groupnorm_apply_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _i64, _i64, _vp, _vp, _vp, _i, _vp, _i64, _i, _i, _i, _vp,
                                       _vp, _i, _i]),
    "univs_patchify_normalize": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                      _vp, _i]),
    "univs_layernorm_multi_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i, C.c_float, _vp, _vp, _i, _vp, _i64, _vp]),
    "univs_layernorm_merge2x2_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, C.c_float, _vp, _i]),
}

_lib = None


class UnivsB200Error(RuntimeError):
    pass


def lib():
    """Loads libunivs_b200.so (built by univs_b200/csrc/build.sh or __graft_entry__.build()).
    Fails loudly if it is missing: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UnivsB200Error(
                f"{LIB_PATH} not found: build it with `bash univs_b200/csrc/build.sh` "
                "(there is no CPU / PyTorch fallback for the hot-path kernels)")
        from . import switches
        switches.export_native()
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().univs_b200_last_error().decode("utf-8", "replace")
        raise UnivsB200Error(f"{what} failed (code {rc}): {msg}")
