"""univs_b200/compat/MultiScaleDeformableAttention.py under the REFERENCE's own caller: ops/functions/ms_deform_attn_func.py
is loaded by path with `MultiScaleDeformableAttention` resolving to the compat module (what putting univs_b200/compat on
PYTHONPATH does), and `MSDeformAttnFunction.apply` is called the way ops/modules/ms_deform_attn.py:120 calls it.
No GPU here: CUDA tensors are imitated and the library is a recorder, so what is checked is the binding (argument order
and types against the header's signature, output allocation, error translation)."""
import ctypes
import importlib.util
import os
import sys

import pytest
import torch

from oracle import ref_shim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_compat():
    spec = importlib.util.spec_from_file_location("MultiScaleDeformableAttention",
                                                  os.path.join(ROOT, "univs_b200", "compat", "MultiScaleDeformableAttention.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Recorder:
    def __init__(self, rc=0):
        self.calls, self.rc = [], rc

    def univs_ms_deform_attn_forward_f32(self, *args):
        self.calls.append(args)
        return self.rc

    def univs_b200_last_error(self):
        return b"ms_deform_attn_forward: spatial_shapes do not cover the value tensor"


class _CudaLike(torch.Tensor):
    """a CPU tensor that reports is_cuda (the binding only reads metadata and data_ptr)"""

    @staticmethod
    def __new__(cls, t):
        return torch.Tensor._make_subclass(cls, t)

    is_cuda = property(lambda self: True)

    def new_empty(self, shape):
        return torch.empty(shape, dtype=self.dtype)


def _inputs(cuda=True):
    wrap = _CudaLike if cuda else (lambda t: t)
    N, M, D, Lq, L, P = 2, 8, 32, 5, 3, 4
    shapes = torch.tensor([(4, 6), (2, 3), (1, 2)], dtype=torch.int64)
    lsi = torch.tensor([0, 24, 30], dtype=torch.int64)
    S = 32
    return (wrap(torch.rand(N, S, M, D)), wrap(shapes), wrap(lsi), wrap(torch.rand(N, Lq, M, L, P, 2)),
            wrap(torch.rand(N, Lq, M, L, P))), (N, S, M, D, L, Lq, P)


def test_binding_argument_order_and_errors(monkeypatch):
    mod = _load_compat()
    rec = _Recorder()
    monkeypatch.setattr(mod, "_lib", rec)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: type("S", (), {"cuda_stream": 77})())
    (value, shapes, lsi, loc, w), dims = _inputs()
    out = mod.ms_deform_attn_forward(value, shapes, lsi, loc, w, 64)
    assert out.shape == (2, 5, 8 * 32) and out.dtype == torch.float32
    (args,) = rec.calls
    assert args[0] == 77 and args[1:6] == (value.data_ptr(), shapes.data_ptr(), lsi.data_ptr(), loc.data_ptr(), w.data_ptr())
    assert args[6:13] == dims and args[13] == out.data_ptr()
    # the header declares exactly this parameter list
    header = open(os.path.join(ROOT, "include", "univs_b200.h")).read()
    decl = header[header.index("int univs_ms_deform_attn_forward_f32("):]
    decl = decl[:decl.index(";")]
    assert decl.count(",") + 1 == len(args) == 14
    # violated preconditions -> RuntimeError, like AT_ASSERTM in ms_deform_attn_cuda.cu:33-43
    (cpu_value, *_), _ = _inputs(cuda=False)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        mod.ms_deform_attn_forward(cpu_value, shapes, lsi, loc, w, 64)
    with pytest.raises(RuntimeError, match="contiguous"):
        mod.ms_deform_attn_forward(_CudaLike(torch.rand(2, 32, 32, 8).transpose(2, 3)), shapes, lsi, loc, w, 64)
    with pytest.raises(RuntimeError, match="int64"):
        mod.ms_deform_attn_forward(value, _CudaLike(shapes.int()), lsi, loc, w, 64)
    monkeypatch.setattr(mod, "_lib", _Recorder(rc=-1))
    with pytest.raises(RuntimeError, match="do not cover"):
        mod.ms_deform_attn_forward(value, shapes, lsi, loc, w, 64)
    with pytest.raises(RuntimeError, match="inference-only"):
        mod.ms_deform_attn_backward(value, shapes, lsi, loc, w, out, 64)


def test_real_library_exports_what_the_stub_binds():
    lib = ctypes.CDLL(os.path.join(ROOT, "univs_b200", "lib", "libunivs_b200.so"))
    assert hasattr(lib, "univs_ms_deform_attn_forward_f32") and hasattr(lib, "univs_b200_last_error")


@pytest.mark.reference
def test_reference_function_runs_on_the_compat_module(monkeypatch):
    """ops/functions/ms_deform_attn_func.py, unmodified, with `import MultiScaleDeformableAttention as MSDA` -> compat"""
    mod = _load_compat()
    rec = _Recorder()
    monkeypatch.setattr(mod, "_lib", rec)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: type("S", (), {"cuda_stream": 5})())
    monkeypatch.setitem(sys.modules, "MultiScaleDeformableAttention", mod)
    path = os.path.join(ref_shim.REF_ROOT, "mask2former/modeling/pixel_decoder/ops/functions/ms_deform_attn_func.py")
    spec = importlib.util.spec_from_file_location("_ref_msda_func", path)
    func = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(func)
    assert func.MSDA is mod
    (value, shapes, lsi, loc, w), dims = _inputs()
    out = func.MSDeformAttnFunction.apply(value, shapes, lsi, loc, w, 128)        # ms_deform_attn.py:120
    assert out.shape == (2, 5, 256) and len(rec.calls) == 1 and rec.calls[0][6:13] == dims


@pytest.mark.gpu
def test_compat_module_on_cuda_matches_the_oracle():
    from oracle import ops_ref
    mod = _load_compat()
    torch.manual_seed(3)                                         # shapes / seed of the reference's ops/test.py:24-33
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    w = torch.rand(N, Lq, M, L, P) + 1e-5
    w = w / w.sum(-1, keepdim=True).sum(-2, keepdim=True)
    got = mod.ms_deform_attn_forward(value.cuda(), shapes.cuda(), lsi.cuda(), loc.cuda(), w.cuda(), 2)
    want = ops_ref.ms_deform_attn(value, shapes.tolist(), lsi.tolist(), loc, w)
    assert (got.cpu() - want).abs().max().item() <= 1e-5 * max(want.abs().max().item(), 1e-6) + 1e-9
