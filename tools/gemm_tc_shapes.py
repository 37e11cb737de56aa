"""Micro-benchmark of csrc/gemm_tc.cu epilogue variants at chosen shapes (run on the B200 box):
  python tools/gemm_tc_shapes.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from univs_b200 import ops, switches  # noqa: E402

switches.export_native()


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
    return a.elapsed_time(b) / n * 1e3


torch.manual_seed(0)
for (M, N, K) in [(294400, 192, 192), (294400, 256, 192), (294400, 128, 192), (294400, 192, 768), (294400, 576, 192), (73600, 384, 384),
                  (18400, 768, 768), (18400, 3072, 768), (18400, 768, 1536)]:
    x = ops.split_operand(torch.randn(M, K, device="cuda"), "f16c")
    w = ops.split_operand(torch.randn(N, K, device="cuda") * 0.05, "f16c")
    bias = torch.randn(N, device="cuda")
    add = torch.randn(M, N, device="cuda")
    out = torch.empty(M, N, device="cuda")
    offs = (0, K)
    t_plain = timeit(lambda: ops.gemm_f16x3_tc(x, offs, w, offs, K, 1.0, bias, None, out=out))
    t_add = timeit(lambda: ops.gemm_f16x3_tc(x, offs, w, offs, K, 1.0, bias, add, out=out))
    t_inpl = timeit(lambda: ops.gemm_f16x3_tc(x, offs, w, offs, K, 1.0, bias, out, out=out))
    t_gelu = timeit(lambda: ops.gemm_f16x3_tc(x, offs, w, offs, K, 1.0, bias, None, want_f32=False, want_operand=True, act=1))
    t_relu = timeit(lambda: ops.gemm_f16x3_tc(x, offs, w, offs, K, 1.0, bias, None, want_f32=False, want_operand=True, act=2))
    fl = 3 * 2 * M * N * K
    print(f"M={M} N={N} K={K}: plain {t_plain:.0f} us ({fl / t_plain / 1e6:.0f} TF/s)  +addend {t_add:.0f}  in-place {t_inpl:.0f}  "
          f"gelu->operand {t_gelu:.0f}  relu->operand {t_relu:.0f}")
print("ok")
