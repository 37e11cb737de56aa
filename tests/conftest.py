import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree at /root/reference")


def _emulate_gpu():
    """UNIVS_EMULATE=1: run `-m gpu` tests WITHOUT a GPU on the CPU emulator of tests/emu (the kernels' sources compiled by
    g++): `.cuda()` hands out emulator-backed tensors, the C ABI is the emulated library.  For checking the gated GPU tests
    themselves (shapes, calls, tolerances) before they get GPU time; big cases are only feasible on the hardware."""
    import torch
    from tests.emu import build_emu
    from tests.test_kernels_cpu_emulation import _Dev, _load, plain
    from univs_b200 import _cabi, ops
    as_dev = lambda t: t if isinstance(t, _Dev) else torch.Tensor._make_subclass(_Dev, t)
    _cabi._lib = _load(build_emu.build())
    ops._stream = lambda: 0
    chk = ops._chk
    ops._chk = lambda t, name, dtype=torch.float32: chk(as_dev(t), name, dtype)
    ops._chk_t = lambda t, name, dtype=torch.float32: (chk(as_dev(t), name, dtype), as_dev(t))[1]
    ops._on_device = lambda t: True
    torch.cuda.is_available = lambda: True
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.set_device = lambda *a, **k: None
    torch.Tensor.cuda = lambda self, *a, **k: as_dev(self)
    real_to = torch.Tensor.to
    is_cuda = lambda d: (isinstance(d, str) and d.startswith("cuda")) or (isinstance(d, torch.device) and d.type == "cuda")

    def to(self, *a, **k):
        if not (any(is_cuda(x) for x in a) or is_cuda(k.get("device"))):
            return real_to(self, *a, **k)
        dtype = next((x for x in a if isinstance(x, torch.dtype)), k.get("dtype"))      # .to("cuda", dtype, non_blocking)
        return as_dev(real_to(self, dtype) if dtype is not None else self)
    torch.Tensor.to = to
    torch.Tensor.cpu = lambda self, *a, **k: plain(self)
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.Tensor.is_cuda = property(lambda self: True)       # intermediates of a model forward are plain CPU tensors
    real_addmm = torch.addmm

    def addmm(inp, a, b, *, beta=1, alpha=1, out_dtype=None, out=None):
        """the library's fp16 x fp16 -> fp32 GEMM (CUDA only) for the fp16x3 policy: exact products, fp32 accumulation"""
        if out_dtype is None:
            return real_addmm(inp, a, b, beta=beta, alpha=alpha) if out is None else real_addmm(inp, a, b, beta=beta, alpha=alpha, out=out)
        r = alpha * (plain(a).float() @ plain(b).float())
        if beta != 0:
            r = r + beta * plain(inp).float()
        if out is not None:
            out.copy_(r)
            return out
        return r
    torch.addmm = addmm


def pytest_collection_modifyitems(config, items):
    import torch
    if os.environ.get("UNIVS_EMULATE") == "1" and not torch.cuda.is_available():
        _emulate_gpu()
    has_gpu = torch.cuda.is_available()
    from oracle import ref_shim
    has_ref = ref_shim.available()
    for it in items:
        if "gpu" in it.keywords and not has_gpu:
            it.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in it.keywords and not has_ref:
            it.add_marker(pytest.mark.skip(reason="reference tree absent (GPU box)"))
