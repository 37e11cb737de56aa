"""Tile-shape sweep of the tiled MSDeformAttn encoder kernel at the north-star shape (run on the B200 box):
  python tools/msda_tile_sweep.py
Checks every tile width bit-for-bit against the untiled kernel (same per-(frame, query, head) arithmetic) and prints
CUDA-event timings; the L1 / L2 hit rates behind the numbers come from `ncu -k regex:msda_encoder`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from univs_b200 import ops          # noqa: E402

shapes = [(23, 40), (46, 80), (92, 160)]
starts = [0, 920, 920 + 3680]
S = sum(h * w for h, w in shapes)
torch.manual_seed(0)
val = torch.randn(5, S, 8, 32, device="cuda")
ol = torch.cat([torch.randn(5, S, 192, device="cuda") * 2.0, torch.randn(5, S, 96, device="cuda")], -1).contiguous()


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


ref = ops.ms_deform_attn_encoder(val, shapes, starts, ol, tile=0)
print(f"untiled: {timeit(lambda: ops.ms_deform_attn_encoder(val, shapes, starts, ol, tile=0)) * 1e3:.1f} us")
for tile in (1, 2, 4, 8, 16, 32):
    out = ops.ms_deform_attn_encoder(val, shapes, starts, ol, tile=tile)
    same = torch.equal(out, ref)
    t = timeit(lambda: ops.ms_deform_attn_encoder(val, shapes, starts, ol, tile=tile))
    print(f"tile {tile:2d} x {32 // tile:2d}: {t * 1e3:.1f} us  bit-identical={same}")
