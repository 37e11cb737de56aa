"""csrc/gemm_tc.cu: one-CTA kernel against the CTA-pair (cta_group::2) variant (UNIVS_GEMM_PAIR) at the Swin-L layer shapes of
the north-star clip -- bit equality of the results and time per launch (run on the B200 box):
  python tools/gemm_tc_pair.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from univs_b200 import ops, switches  # noqa: E402

switches.export_native()


def timeit(fn, n=20):
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
SHAPES = [(300, 260, 192), (73600, 1152, 384), (73600, 384, 384), (73600, 1536, 384), (73600, 384, 1536),
          (18400, 2304, 768), (18400, 768, 768), (18400, 3072, 768), (18400, 768, 1536),
          (4600, 4608, 1536), (4600, 1536, 1536), (4600, 6144, 1536), (20000, 256, 1024), (18400, 900, 768)]
tot = {"0": 0.0, "1": 0.0}
for (M, N, K) in SHAPES:
    x = ops.split_operand(torch.randn(M, K, device="cuda"), "f16c")
    w = ops.split_operand(torch.randn(N, K, device="cuda") * 0.05, "f16c")
    bias = torch.randn(N, device="cuda")
    add = torch.randn(M, N, device="cuda")
    offs = (0, K)
    res, t = {}, {}
    for mode in ("0", "1"):
        os.environ["UNIVS_GEMM_PAIR"] = mode
        res[mode] = ops.gemm_f16x3_tc(x, offs, w, offs, K, 1.0, bias, add, want_f32=True, want_operand=True, act=1)
        torch.cuda.synchronize()
        out = torch.empty(M, N, device="cuda")
        t[mode] = timeit(lambda: ops.gemm_f16x3_tc(x, offs, w, offs, K, 1.0, bias, None, out=out))
        tg = timeit(lambda: ops.gemm_f16x3_tc(x, offs, w, offs, K, 1.0, bias, None, want_f32=False, want_operand=True, act=1))
        t[mode + "g"] = tg
        if M > 1000:
            tot[mode] += t[mode]
    same = torch.equal(res["0"][0], res["1"][0]) and torch.equal(res["0"][1], res["1"][1])
    fl = 3 * 2 * M * N * K
    print(f"M={M} N={N} K={K}: one-CTA {t['0']:.0f} us ({fl / t['0'] / 1e6:.0f} TF/s)  cluster {t['1']:.0f} us "
          f"({fl / t['1'] / 1e6:.0f} TF/s)  gelu->operand {t['0g']:.0f} / {t['1g']:.0f}  bit-equal {same}", flush=True)
    assert same
os.environ.pop("UNIVS_GEMM_PAIR")
print(f"sum one-CTA {tot['0']:.0f} us, cluster {tot['1']:.0f} us")
print("ok")
