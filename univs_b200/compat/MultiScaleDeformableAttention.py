"""Drop-in for the compiled extension of mask2former/modeling/pixel_decoder/ops/setup.py: the Python-visible module
`MultiScaleDeformableAttention` with `ms_deform_attn_forward` / `ms_deform_attn_backward` (ops/src/vision.cpp:18-21),
bound to `univs_ms_deform_attn_forward_f32` of libunivs_b200.so (include/univs_b200.h) over ctypes -- no torch types cross
the boundary.  Same contract as ms_deform_attn_cuda_forward (ops/src/cuda/ms_deform_attn_cuda.cu:25-85): contiguous CUDA
fp32 inputs, int64 CUDA level tables, a freshly allocated output [N, Lq, M*D] owned by the caller, work queued on the
current stream without synchronisation, RuntimeError on a violated precondition.  `im2col_step` is accepted and ignored
(the reference uses it only to chunk the batch, :55-80).

Use: put `univs_b200/compat` on PYTHONPATH (or copy this file next to the reference's ops/functions/); the reference's
`MSDeformAttnFunction.forward` (ms_deform_attn_func.py:34-39) then runs this operator unmodified."""
import ctypes
import os

import torch

_LIB_PATH = os.environ.get("UNIVS_B200_LIB",
                           os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lib", "libunivs_b200.so"))
_lib = None


def _library():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(_LIB_PATH)          # fails loudly when the library has not been built: there is no fallback
        lib.univs_b200_last_error.restype = ctypes.c_char_p
        f = lib.univs_ms_deform_attn_forward_f32
        f.restype = ctypes.c_int
        f.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 7 + [ctypes.c_void_p]
        _lib = lib
    return _lib


def _require(t, name, dtype):
    # ms_deform_attn_cuda.cu:33-43 (AT_ASSERTM ... must be contiguous / must be a CUDA tensor)
    if not t.is_contiguous():
        raise RuntimeError(f"{name} tensor has to be contiguous")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype} (this build computes in float32), got {t.dtype}")


def ms_deform_attn_forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, im2col_step):
    _require(value, "value", torch.float32)
    _require(spatial_shapes, "spatial_shapes", torch.int64)
    _require(level_start_index, "level_start_index", torch.int64)
    _require(sampling_loc, "sampling_loc", torch.float32)
    _require(attn_weight, "attn_weight", torch.float32)
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_loc.shape
    out = value.new_empty((N, Lq, M * D))
    lib = _library()
    rc = lib.univs_ms_deform_attn_forward_f32(
        torch.cuda.current_stream().cuda_stream, value.data_ptr(), spatial_shapes.data_ptr(), level_start_index.data_ptr(),
        sampling_loc.data_ptr(), attn_weight.data_ptr(), N, S, M, D, L, Lq, P, out.data_ptr())
    if rc != 0:
        raise RuntimeError(lib.univs_b200_last_error().decode("utf-8", "replace"))
    return out


def ms_deform_attn_backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_output, im2col_step):
    raise RuntimeError("ms_deform_attn_backward: this is an inference-only build (univs_ms_deform_attn_backward_f32 "
                       "returns UNIVS_E_NOTIMPL)")
