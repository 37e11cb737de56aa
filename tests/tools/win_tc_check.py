"""Staged diagnostics + timing of the tcgen05 window-attention kernel (run on the B200 box):
  python tests/tools/win_tc_check.py [--time]
Stage 1 compares the biased, masked scores (QK^T path: loader, swizzled operand tiles, K-major descriptors),
stage 2 the attention output (softmax, P tile, PV path with the MN-major V descriptor), stage 3 the GEMM-operand output,
each against the CPU oracle and the validated mma.sync kernel; --time adds CUDA-event timings at the four
north-star Swin-L stage shapes (T=5, 736x1280) next to the mma.sync kernel."""
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


def check(B, H, W, nH, shift, ver=0):
    torch.manual_seed(1)
    C = 32 * nH
    qkv, bias, table = torch.randn(B, H, W, 3 * C), torch.randn(3 * C) * 0.3, torch.randn(529, nH) * 0.5
    want, scores = ops_ref.swin_window_attention(qkv, bias, table, nH, 12, shift, return_scores=True)
    out, op, dbg = ops.swin_window_attention_tc(qkv.cuda(), bias.cuda(), table.cuda(), nH, shift, True, True, ver, True)
    torch.cuda.synchronize()
    dbg = dbg.view(scores.shape).cpu()
    e_s = rel(dbg, scores)
    # where do the scores differ: per row tile (rows < 128 vs tail) and per column half
    d = (dbg - scores).abs().amax(dim=(0, 1, 2))
    parts = {"tile0/half0": d[:128, :72].max().item(), "tile0/half1": d[:128, 72:].max().item(),
             "tail/keys0-63": d[128:, :64].max().item(), "tail/keys64-143": d[128:, 64:].max().item()}
    e_o = rel(out, want)
    opf = op.float().cpu()
    e_op = rel(opf[..., 2 * C:] + opf[..., :C] * 2.0 ** -11, want)
    ref = ops.swin_window_attention(qkv.cuda(), bias.cuda(), table.cuda(), nH, 12, shift, precision=0)
    print(f"B={B} {H}x{W} heads={nH} shift={shift}: scores {e_s:.2e} out {e_o:.2e} operand {e_op:.2e} "
          f"vs mma.sync {rel(out, ref):.2e}  score abs-diff by part {parts}")
    return max(e_s, e_o, e_op) < 2e-5


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
    ap.add_argument("--ver", type=int, default=0, help="kernel version (flags bit 0)")
    a = ap.parse_args()
    ok = True
    for shp in [(1, 12, 12, 1, 0), (1, 24, 36, 2, 0), (2, 24, 27, 4, 6), (1, 46, 80, 6, 6), (3, 23, 40, 24, 6), (2, 36, 48, 1, 6)]:
        ok &= check(*shp, ver=a.ver)
    print("PARITY", "ok" if ok else "FAILED")
    if a.time:
        for (H, W, nH) in [(184, 320, 6), (92, 160, 12), (46, 80, 24), (23, 40, 48)]:
            C = 32 * nH
            qkv = torch.randn(5, H, W, 3 * C, device="cuda")
            bias, table = torch.randn(3 * C, device="cuda"), torch.randn(529, nH, device="cuda")
            for shift in (0, 6):
                t_tc = timeit(lambda: ops.swin_window_attention_tc(qkv, bias, table, nH, shift, False, True, a.ver, compact=bool(a.ver)))
                t_mma = timeit(lambda: ops.swin_window_attention_operand(qkv, bias, table, nH, 12, shift))
                gb = qkv.numel() * 4 + qkv.numel() // 3 * 6
                print(f"stage {H}x{W} heads={nH} shift={shift}: tcgen05 {t_tc:.3f} ms ({gb / t_tc / 1e6:.0f} GB/s)  "
                      f"mma.sync {t_mma:.3f} ms ({gb / t_mma / 1e6:.0f} GB/s)")
