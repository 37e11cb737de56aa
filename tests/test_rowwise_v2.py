"""8-wide GELU / ReLU / operand-split kernel (csrc/elementwise.cu gelu_split8_kernel, UNIVS_ROWWISE_V2 bit 0) and the
wide-store LayerNorm (layernorm_wide_kernel, bit 1): must be BIT-identical to the validated gelu_split_kernel /
layernorm_kernel (same erff, same reductions, same conversions).  The switch is read once per process, so
the two variants run in two child processes and their outputs are compared byte for byte.
Opt-in until it has run on a B200: UNIVS_GPU_ROWWISE_V2=1."""
import os
import subprocess
import sys
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_CHILD = r"""
import sys, torch
sys.path.insert(0, sys.argv[2])
from univs_b200 import ops
torch.manual_seed(0)
out = {}
for rows, C in [(1, 8), (37, 48), (1000, 192), (4600, 768), (333, 6144), (7, 2048), (30001, 384)]:
    x = (torch.randn(rows, C, device="cuda") * 3).contiguous()
    x[0, 0] = 70000.0          # saturates the fp16 hi part
    b = torch.randn(C, device="cuda")
    for fmt in ("f16", "f16u", "f16c"):
        out[f"gelu_{rows}_{C}_{fmt}"] = ops.gelu(x, split=fmt, bias=b).cpu()
        out[f"relu_{rows}_{C}_{fmt}"] = ops.relu(x, split=fmt, bias=None).cpu()
        out[f"split_{rows}_{C}_{fmt}"] = ops.split_operand(x, fmt).cpu()
        if C <= 4096:                                       # LayerNorm (+ residual + deferred bias), fp32 sum and operand
            w, bb, r = torch.randn(C, device="cuda"), torch.randn(C, device="cuda"), torch.randn(rows, C, device="cuda")
            ssum, y = ops.layernorm(x, w, bb, 1e-5, r, True, fmt, b)
            out[f"ln_{rows}_{C}_{fmt}"] = y.cpu()
            out[f"lnsum_{rows}_{C}_{fmt}"] = ssum.cpu()
            out[f"ln_plain_{rows}_{C}_{fmt}"] = ops.layernorm(x, w, bb, 1e-5, None, False, fmt, None)[1].cpu()
torch.cuda.synchronize()
torch.save(out, sys.argv[1])
"""


@pytest.mark.gpu
def test_rowwise_v2_bit_identical():
    with tempfile.TemporaryDirectory() as d:
        res = {}
        for v2 in ("0", "3", "7"):     # 3 = bit 0 (GELU / ReLU / split) + bit 1 (wide-store LayerNorm); 7 = + bit 2 (streaming LayerNorm)
            path = os.path.join(d, f"v{v2}.pt")
            env = dict(os.environ, UNIVS_ROWWISE_V2=v2)
            subprocess.run([sys.executable, "-c", _CHILD, path, ROOT], check=True, env=env, timeout=600)
            res[v2] = torch.load(path)
    assert res["0"].keys() == res["3"].keys() == res["7"].keys() and len(res["0"]) == 63 + 54
    for v2 in ("3", "7"):
        for k in res["0"]:
            a, b = res["0"][k], res[v2][k]
            assert a.dtype == b.dtype and a.shape == b.shape, k
            assert torch.equal(a.contiguous().view(torch.int16), b.contiguous().view(torch.int16)), (v2, k)
