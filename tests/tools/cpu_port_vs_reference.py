"""Calibration of bench.py's CPU arm (kind "port": product host logic + oracle/ops_ref.py operators) against the REFERENCE'S
OWN FILES loaded by path (oracle/ref_shim.py; needs /root/reference, i.e. the development container): same weights, same
clip, same thread count, fp32.  BASELINE config C2 geometry (Swin-T, 480x864, Q=100) at T frames (default 2).
  python tests/tools/cpu_port_vs_reference.py [T]        -> one JSON line (recorded in BASELINE.md)"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import ref_shim                      # noqa: E402
from oracle.cpu_backend import oracle_ops        # noqa: E402
from tests import model_factory as mf            # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 2
H, W, Q = 480, 864, 100
swin = dict(embed_dim=96, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=7)
torch.manual_seed(0)
clip = mf.make_clip_emb()
ref = ref_shim.build_reference_model(swin, num_queries=Q, num_frames=T, clip_emb=clip)
prod = mf.build_product_model(swin, num_queries=Q, num_frames=T, clip_emb=clip)
for r, p in zip(ref, prod):
    sd = mf.keyed_state_dict(r.state_dict())
    r.load_state_dict(sd)
    p.load_state_dict(sd)
x = torch.randn(T, 3, H, W)
tg = lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual", "frame_indices": torch.arange(T)}]
threads = torch.get_num_threads()


def best(fn, n=3):
    fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


with torch.no_grad():
    t_ref = best(lambda: ref_shim.reference_clip_forward(*ref, x, tg()))
    with oracle_ops():
        t_port = best(lambda: mf.product_clip_forward(*prod, x, tg()))
        rout = ref_shim.reference_clip_forward(*ref, x, tg())[2]
        pout = mf.product_clip_forward(*prod, x, tg())[2]
err = (pout["pred_masks"] - rout["pred_masks"]).abs().max().item() / rout["pred_masks"].abs().max().item()
print(json.dumps({"config": f"C2 geometry: Swin-T T={T} {H}x{W} Q={Q}, fp32, {threads} threads", "reference_by_path_s": t_ref,
                  "port_s": t_port, "port_over_reference_time": t_port / t_ref, "reference_frames_per_s": T / t_ref,
                  "port_frames_per_s": T / t_port, "pred_masks_rel_diff": err}))
