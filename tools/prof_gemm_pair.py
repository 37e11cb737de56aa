"""One large launch of the dense-layer kernel per variant at Swin-L stage-3 shapes, for `ncu --set full`:
  python tools/prof_gemm_pair.py      (launch order: fc1+GELU->operand pair, the same one-CTA, qkv pair, fc2 slice + addend pair)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from univs_b200 import ops, switches  # noqa: E402

switches.export_native()
torch.manual_seed(0)


def operands(M, N, K):
    x = ops.split_operand(torch.randn(M, K, device="cuda"), "f16c")
    w = ops.split_operand(torch.randn(N, K, device="cuda") * 0.05, "f16c")
    return x, w, torch.randn(N, device="cuda")


x, w, b = operands(18400, 3072, 768)
for mode in ("1", "0"):
    os.environ["UNIVS_GEMM_PAIR"] = mode
    ops.gemm_f16x3_tc(x, (0, 768), w, (0, 768), 768, 1.0, b, None, want_f32=False, want_operand=True, act=1)
os.environ["UNIVS_GEMM_PAIR"] = "1"
x, w, b = operands(18400, 2304, 768)
ops.gemm_f16x3_tc(x, (0, 768), w, (0, 768), 768, 1.0, b, None)
x, w, b = operands(18400, 768, 1536)
out = torch.randn(18400, 768, device="cuda")
ops.gemm_f16x3_tc(x, (0, 1536), w, (0, 1536), 1536, 1.0, b, out, out=out)
torch.cuda.synchronize()
print("ok")
