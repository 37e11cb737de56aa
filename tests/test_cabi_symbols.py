"""CPU-side checks of the drop-in boundary: the shared library loads and exports exactly the entry points that
include/univs_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "univs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(univs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from univs_b200 import _cabi
    if not os.path.exists(_cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/univs_b200.h but not exported"
        assert n in _cabi.SIGNATURES, f"{n} has no ctypes signature in univs_b200/_cabi.py"
    assert sorted(_cabi.SIGNATURES) == names
    assert _cabi.lib().univs_b200_abi_version() == 1


def test_backward_is_exported_but_not_implemented():
    from univs_b200 import _cabi
    rc = _cabi.lib().univs_ms_deform_attn_backward_f32()
    assert rc == -3
    assert b"out of scope" in _cabi.lib().univs_b200_last_error()


def test_ops_refuse_cpu_tensors():
    import torch
    from univs_b200 import ops, _cabi
    with pytest.raises(_cabi.UnivsB200Error):
        ops.mask_einsum(torch.zeros(1, 4, 32), torch.zeros(1, 8, 32))


def test_new_entry_points_validate_arguments_before_touching_the_gpu():
    """argument errors are reported through the return code + univs_b200_last_error without any CUDA call"""
    from univs_b200 import _cabi
    L = _cabi.lib()
    one = ctypes.c_void_p(16)          # any non-null pointer: validation fails before it is dereferenced
    err = lambda: L.univs_b200_last_error().decode()
    # tcgen05 window attention: 12x12 windows only, head_dim 32, shift < window, no flags
    assert L.univs_swin_window_attention_tc(None, one, one, one, 1, 24, 24, 64, 2, 7, 0, 0, one, None, None) == -1
    assert "12x12" in err()
    assert L.univs_swin_window_attention_tc(None, one, one, one, 1, 24, 24, 60, 2, 12, 0, 0, one, None, None) == -1
    assert "head_dim" in err()
    assert L.univs_swin_window_attention_tc(None, one, one, one, 1, 24, 24, 64, 2, 12, 12, 0, one, None, None) == -1
    assert L.univs_swin_window_attention_tc(None, one, one, one, 1, 24, 24, 64, 2, 12, 6, 2, one, None, None) == -1
    assert L.univs_swin_window_attention_tc(None, one, one, one, 1, 24, 24, 64, 2, 12, 6, 0, None, None, None) == -1
    assert L.univs_swin_window_attention_tc(None, one, one, one, 0, 24, 24, 64, 2, 12, 6, 0, one, None, None) == 0     # empty batch
    # tcgen05 cross-attention: at most 256 queries, channels = heads * 32
    assert L.univs_mha_tc_forward_f32(None, one, one, one, None, None, 0, 1, 300, 1000, 256, 0, one, one) == -1
    assert "256 queries" in err()
    assert L.univs_mha_tc_forward_f32(None, one, one, one, None, None, 0, 1, 100, 1000, 250, 0, one, one) == -1
    assert L.univs_mha_tc_forward_f32(None, one, one, one, None, None, 0, 1, 100, 0, 256, 0, one, one) == -1
    assert L.univs_mha_tc_forward_f32(None, one, one, one, None, None, 0, 1, 100, 1000, 256, 4, one, one) == -1
    assert L.univs_mha_tc_forward_f32(None, one, one, one, None, None, 0, 0, 100, 1000, 256, 0, one, one) == 0
    assert L.univs_mha_tc_workspace_bytes(5, 200, 14720, 256) > 16 and L.univs_mha_tc_workspace_bytes(0, 200, 100, 256) == 0
    # pooled mask features: even integer ratios only
    assert L.univs_mask_feature_pool_f32(None, one, 1, 16, 24, 32, 5, 6, one, 0) == -1
    assert L.univs_mask_feature_pool_f32(None, one, 1, 16, 24, 32, 16, 24, one, 0) == -1
    assert "even" in err()
    assert L.univs_mask_feature_pool_f32(None, one, 1, 16, 24, 32, 8, 12, one, 5) == -1
    assert L.univs_attn_mask_bits_direct_f32(None, one, 3, 2, 0, one, one) == -1
    assert L.univs_attn_mask_bits_direct_f32(None, None, 0, 2, 10, None, None) == 0
