"""Timing probe of the tcgen05 window-attention loader (UNIVS_WINTC_PROBE bit 0: no split arithmetic, bit 1: no global loads;
results are garbage with a probe on).  python tools/wintc_probe.py"""
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
for (H, W, nH) in [(184, 320, 6), (92, 160, 12), (46, 80, 24), (23, 40, 48)]:
    C = 32 * nH
    qkv = torch.randn(5, H, W, 3 * C, device="cuda")
    bias, table = torch.randn(3 * C, device="cuda"), torch.randn(529, nH, device="cuda")
    row = []
    for probe in ("0", "1", "2", "3"):
        os.environ["UNIVS_WINTC_PROBE"] = probe
        row.append(timeit(lambda: ops.swin_window_attention_tc(qkv, bias, table, nH, 6, False, True)))
    print(f"{H}x{W} heads {nH}: normal {row[0]:.0f} us | no split math {row[1]:.0f} | no global loads {row[2]:.0f} | neither {row[3]:.0f}",
          flush=True)
os.environ.pop("UNIVS_WINTC_PROBE")
print("ok")
