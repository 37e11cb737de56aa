"""Probe: error of TF32 tensor-core GEMMs (cuBLAS) on exactly-TF32-representable operands vs fp64, as a function of
the K-chunk length -- isolates the accumulation error of the tensor-core datapath (products are exact)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from univs_b200 import ops
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = True
M, N = 8192, 768
for K in (192, 768, 3072):
    x = ops.round_tf32(torch.randn(M, K, device="cuda"))
    w = ops.round_tf32(torch.randn(N, K, device="cuda") * 0.05)
    want = x.double() @ w.double().t()
    sc = want.abs().max().item()
    for chunk in (K, 256, 128, 64, 32):
        if chunk > K: continue
        y = torch.zeros(M, N, device="cuda")
        for k0 in range(0, K, chunk):
            y.addmm_(x[:, k0:k0 + chunk], w[:, k0:k0 + chunk].t())
        err = (y.double() - want)
        print(f"K={K} chunk={chunk}: max {err.abs().max().item()/sc:.2e} mean-signed {(err*want.sign()).mean().item()/want.abs().mean().item():.2e}", flush=True)
    torch.backends.cuda.matmul.allow_tf32 = False
    y = x @ w.t()
    err = (y.double() - want)
    print(f"K={K} fp32 SIMT: max {err.abs().max().item()/sc:.2e} mean-signed {(err*want.sign()).mean().item()/want.abs().mean().item():.2e}", flush=True)
    torch.backends.cuda.matmul.allow_tf32 = True
