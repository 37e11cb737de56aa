"""Timing of csrc/gemm_tc.cu next to the library GEMM path it replaces, at the north-star shapes (run on the B200 box):
  python tools/gemm_tc_check.py
Per shape: the plain dense layer (library: one fp16 GEMM over the 3K-wide operands) and the MLP chain
fc1 -> GELU -> operand -> fc2 (library: GEMM + gelu_split kernel + GEMM)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from univs_b200 import nn_ops, ops, switches  # noqa: E402
from univs_b200.precision import set_precision  # noqa: E402

switches.export_native()
set_precision("fp16x3")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


torch.manual_seed(0)
T = 5
for name, M, C in [("stage1", T * 184 * 320, 192), ("stage2", T * 92 * 160, 384), ("stage3", T * 46 * 80, 768), ("stage4", T * 23 * 40, 1536)]:
    x = torch.randn(M, C, device="cuda")
    h = nn_ops.prep(x)
    fc1, fc2 = torch.nn.Linear(C, 4 * C).cuda(), torch.nn.Linear(4 * C, C).cuda()
    qkv = torch.nn.Linear(C, 3 * C).cuda()
    res = {}
    for tc in (False, True):
        nn_ops.set_gemm_tc(tc)
        res[("qkv", tc)] = timeit(lambda: nn_ops.linear_prepped(h, qkv.weight, None))
        res[("mlp", tc)] = timeit(lambda: nn_ops.mlp(h, fc1, fc2))
    nn_ops.set_gemm_tc(False)
    y0 = nn_ops.mlp(h, fc1, fc2)
    nn_ops.set_gemm_tc(True)
    y1 = nn_ops.mlp(h, fc1, fc2)
    nn_ops.set_gemm_tc(False)
    err = (y1 - y0).abs().max().item() / y0.abs().max().item()
    fl_qkv, fl_mlp = 3 * 2 * M * C * 3 * C, 3 * 2 * M * C * 4 * C * 2
    print(f"{name} M={M} C={C}: qkv lib {res[('qkv', False)]:.3f} ms ({fl_qkv / res[('qkv', False)] / 1e9:.0f} TF/s)  tc {res[('qkv', True)]:.3f} ms "
          f"({fl_qkv / res[('qkv', True)] / 1e9:.0f} TF/s) | mlp lib {res[('mlp', False)]:.3f} ms  tc {res[('mlp', True)]:.3f} ms "
          f"({fl_mlp / res[('mlp', True)] / 1e9:.0f} TF/s fp16-MMA)  mlp diff {err:.1e}")
print("ok")
