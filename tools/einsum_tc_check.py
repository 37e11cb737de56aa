"""Stand-alone check + timing of the tcgen05 mask einsum (run under `timeout` on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from univs_b200 import ops

torch.manual_seed(0)
ok = True
for (T, Q, C, HW) in [(1, 16, 32, 128), (1, 20, 256, 4096), (2, 100, 256, 1620), (3, 200, 256, 3680), (1, 232, 256, 130), (2, 7, 64, 66), (5, 200, 256, 58880)]:
    E = ops.round_tf32(torch.randn(T, Q, C, device="cuda"))
    F = ops.round_tf32(torch.randn(T, HW, C, device="cuda"))
    got = ops.mask_einsum(E, F, mode="tf32")
    torch.cuda.synchronize()
    ref = ops.mask_einsum_mma(E, F, ops.PREC_TF32)
    want = torch.einsum("tqc,tpc->qtp", E.double(), F.double()) if HW < 10000 else None
    e_mma = (got - ref).abs().max().item() / ref.abs().max().item()
    e64 = (got.double() - want).abs().max().item() / want.abs().max().item() if want is not None else float("nan")
    print(f"T={T} Q={Q} C={C} HW={HW}: tc-vs-mma {e_mma:.2e}  tc-vs-fp64 {e64:.2e}", flush=True)
    ok &= e_mma < 1e-5
# timing at the north-star shape
T, Q, C, HW = 5, 200, 256, 58880
E = ops.round_tf32(torch.randn(T, Q, C, device="cuda")); F = ops.round_tf32(torch.randn(T, HW, C, device="cuda"))
out = torch.empty(Q, T, HW, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
F16 = ops.prepare_mask_features(F, "f16x3")
def _cluster(e, f, out):
    ops._einsum_mc = 1
    try:
        return ops.mask_einsum(e, F16, out=out, mode="f16x3")
    finally:
        ops._einsum_mc = 0


variants = [("tc_tf32", lambda e, f, out: ops.mask_einsum(e, f, out=out, mode="tf32")), ("tc_f16x3", lambda e, f, out: ops.mask_einsum(e, F16, out=out, mode="f16x3")), ("mma", lambda e, f, out: ops.mask_einsum_mma(e, f, ops.PREC_TF32, out=out)), ("mma3x", lambda e, f, out: ops.mask_einsum_mma(e, f, ops.PREC_TF32X3, out=out))]
if os.environ.get("EINSUM_MC") == "1":          # opt-in: the cluster / multicast kernel (csrc/mask_einsum_mc.cu)
    variants.insert(2, ("tc_f16x3_cluster", _cluster))
for name, fn in variants:
    for _ in range(3): fn(E, F, out=out)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(E, F, out=out); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    byts = 4.0 * T * (Q * C + C * HW + Q * HW)
    print(f"{name}: {ms:.3f} ms  -> {byts / ms / 1e6:.0f} GB/s algorithmic", flush=True)
print("OK" if ok else "MISMATCH")
