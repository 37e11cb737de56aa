"""Large launches of single kernels at the north-star shapes, for `ncu --set full -k regex:<kernel> -c N`:
  python tools/prof_targets.py win|winmma|mha|gelu|ln [...]
No parity checks (tests/tools/*_check.py do that): every launch here is one worth profiling."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from univs_b200 import ops, switches  # noqa: E402

switches.export_native()
what = sys.argv[1:] or ["win", "winmma", "mha", "gelu", "ln"]
torch.manual_seed(0)
if "win" in what or "winmma" in what:
    H, W, nH = 184, 320, 6
    C = 32 * nH
    qkv = torch.randn(5, H, W, 3 * C, device="cuda")
    bias, table = torch.randn(3 * C, device="cuda"), torch.randn(529, nH, device="cuda")
    for shift in (0, 6):
        if "win" in what:
            ops.swin_window_attention_tc(qkv, bias, table, nH, shift, False, True, compact=True)   # version: UNIVS_WIN_TC
        if "winmma" in what:
            ops.swin_window_attention_operand(qkv, bias, table, nH, 12, shift)
if "mha" in what:
    for Lk in (3680, 14720):
        q = torch.randn(5, 200, 256, device="cuda")
        k, v = torch.randn(5, Lk, 256, device="cuda"), torch.randn(5, Lk, 256, device="cuda")
        mask = torch.rand(5, 200, Lk, device="cuda") < 0.7
        bits, ro = ops.pack_mask_bits(mask), (~mask.all(-1)).to(torch.int32)
        ops.mha_core_tc(q, k, v, bits, ro, flags=0)
if "gelu" in what:
    x = torch.randn(5 * 184 * 320, 768, device="cuda")
    b = torch.randn(768, device="cuda")
    ops.gelu(x, "f16", b)
if "ln" in what:
    x = torch.randn(5 * 184 * 320, 192, device="cuda")
    r = torch.randn_like(x)
    w, b = torch.randn(192, device="cuda"), torch.randn(192, device="cuda")
    ops.layernorm(x, w, b, 1e-5, r, True, "f16", b)
torch.cuda.synchronize()
print("ok")
