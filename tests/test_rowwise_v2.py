"""8-wide GELU / ReLU / operand-split kernel (csrc/elementwise.cu gelu_split8_kernel, opt-in UNIVS_ROWWISE_V2=1): must be
BIT-identical to the validated gelu_split_kernel (same erff, same conversions).  The switch is read once per process, so
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
for rows, C in [(1, 8), (37, 48), (1000, 192), (4600, 768), (333, 6144), (7, 2048)]:
    x = (torch.randn(rows, C, device="cuda") * 3).contiguous()
    x[0, 0] = 70000.0          # saturates the fp16 hi part
    b = torch.randn(C, device="cuda")
    for fmt in ("f16", "f16u"):
        out[f"gelu_{rows}_{C}_{fmt}"] = ops.gelu(x, split=fmt, bias=b).cpu()
        out[f"relu_{rows}_{C}_{fmt}"] = ops.relu(x, split=fmt, bias=None).cpu()
        out[f"split_{rows}_{C}_{fmt}"] = ops.split_operand(x, fmt).cpu()
torch.cuda.synchronize()
torch.save(out, sys.argv[1])
"""


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("UNIVS_GPU_ROWWISE_V2") != "1", reason="opt-in (UNIVS_GPU_ROWWISE_V2=1): not yet run on a B200")
def test_rowwise_v2_bit_identical():
    with tempfile.TemporaryDirectory() as d:
        res = {}
        for v2 in ("0", "1"):
            path = os.path.join(d, f"v{v2}.pt")
            env = dict(os.environ, UNIVS_ROWWISE_V2=v2)
            subprocess.run([sys.executable, "-c", _CHILD, path, ROOT], check=True, env=env, timeout=600)
            res[v2] = torch.load(path)
    assert res["0"].keys() == res["1"].keys() and len(res["0"]) == 36
    for k in res["0"]:
        assert torch.equal(res["0"][k].view(torch.int16), res["1"][k].view(torch.int16)), k
