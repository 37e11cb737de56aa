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
