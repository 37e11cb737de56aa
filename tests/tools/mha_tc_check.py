"""Parity + timing of the tcgen05 cross-attention kernel (run on the B200 box):  python tests/tools/mha_tc_check.py [--time]
Parity against the CPU oracle and the validated mma.sync kernel for both V staging variants (flags 0: MN-major B
descriptor, flags 1: transposed V, K-major); --time adds CUDA-event timings at the three north-star memory sizes
(T=5 frames, Q=200, 8 heads) next to the mma.sync kernel, with the achieved K/V streaming bandwidth."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ops_ref          # noqa: E402  (diagnostic tool: the oracle is the checker)
from univs_b200 import ops          # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def check(B, Lq, Lk, C, flags):
    torch.manual_seed(3)
    q, k, v = torch.randn(B, Lq, C), torch.randn(B, Lk, C), torch.randn(B, Lk, C)
    mask = torch.rand(B, Lq, Lk) < 0.7
    mask[:, 0] = True
    bits, ro = ops.pack_mask_bits(mask).cuda(), (~mask.all(-1)).to(torch.int32).cuda()
    want = ops_ref.mha_core(q, k, v, C // 32, mask.to(torch.uint8), unmask_full_rows=True)
    got = ops.mha_core_tc(q.cuda(), k.cuda(), v.cuda(), bits, ro, flags=flags)
    ref = ops.mha_core(q.cuda(), k.cuda(), v.cuda(), bits, ro, precision=0)
    torch.cuda.synchronize()
    e = rel(got, want)
    # which query rows are off: tile 0 (rows < 128) vs tile 1
    d = (got.cpu() - want).abs().amax(dim=(0, 2))
    print(f"B={B} Lq={Lq} Lk={Lk} C={C} flags={flags}: vs oracle {e:.2e}  vs mma.sync {rel(got, ref):.2e}  "
          f"rows<128 {d[:128].max().item():.2e}  rows>=128 {(d[128:].max().item() if Lq > 128 else 0.0):.2e}")
    return e < 2e-5


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
    return a.elapsed_time(b) / n


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    a = ap.parse_args()
    ok = True
    for flags in (0, 1):
        for shp in [(1, 7, 128, 32), (2, 200, 920, 256), (3, 256, 130, 256), (5, 200, 3680, 256)]:
            ok &= check(*shp, flags)
    print("PARITY", "ok" if ok else "FAILED")
    if a.time:
        for Lk in (920, 3680, 14720):
            q = torch.randn(5, 200, 256, device="cuda")
            k, v = torch.randn(5, Lk, 256, device="cuda"), torch.randn(5, Lk, 256, device="cuda")
            mask = torch.rand(5, 200, Lk, device="cuda") < 0.7
            bits, ro = ops.pack_mask_bits(mask), (~mask.all(-1)).to(torch.int32)
            t_tc = timeit(lambda: ops.mha_core_tc(q, k, v, bits, ro, flags=0))
            t_mma = timeit(lambda: ops.mha_core(q, k, v, bits, ro, precision=0))
            gb = 2 * k.numel() * 4
            print(f"S={Lk}: tcgen05 {t_tc * 1e3:.1f} us ({gb / t_tc / 1e6:.0f} GB/s of K/V)  mma.sync {t_mma * 1e3:.1f} us "
                  f"({gb / t_mma / 1e6:.0f} GB/s)")
