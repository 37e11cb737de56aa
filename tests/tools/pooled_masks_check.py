"""CPU check of the pooled-feature attention masks (UNIVS_POOLED_MASKS, csrc/decoder_glue.cu) at the north-star geometry
(Swin-L, 720x1280, Q=200, bench initialisation, one frame): product host model with the CPU oracle operators, default path
vs pooled path; counts the attention-mask bits that differ in every intermediate head and compares the outputs."""
import json, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.cpu_backend import oracle_ops
from univs_b200.build import build_model, make_cfg
torch.set_num_threads(8)
g = torch.Generator().manual_seed(0)
cfg = make_cfg("large", 200, 1, clip_emb=torch.randn(3938, 640, generator=g), TEXT_PROMPT_TO_IMAGE_ENABLE=False)
model = build_model(cfg)
frames = torch.rand(1, 3, 720, 1280, generator=g) * 255
tg = lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual", "frame_indices": torch.arange(1)}]
dec = model.sem_seg_head.predictor
res = {}
for pooled in (False, True):
    dec.pooled_masks = pooled
    seen = []
    dec.attn_mask_hook = lambda i, bits, ro: (seen.append((bits.clone(), ro.clone())), (bits, ro))[1]
    t0 = time.time()
    with oracle_ops():
        out = model.clip_forward(frames, tg())
    res[pooled] = (out, seen)
    print("pooled", pooled, "time", time.time() - t0, flush=True)
(o0, s0), (o1, s1) = res[False], res[True]
tot = flips = 0
for i, ((b0, r0), (b1, r1)) in enumerate(zip(s0, s1)):
    x = (b0 ^ b1).to(torch.int64) & 0xFFFFFFFF
    n = sum(int(((x >> k) & 1).sum()) for k in range(32))
    flips += n; tot += b0.numel() * 32
    print("head", i, "flipped bits", n, "of", b0.numel() * 32, "row flags differ", int((r0 != r1).sum()))
pm0, pm1 = o0["pred_masks"], o1["pred_masks"]
print(json.dumps({"workload": "Swin-L 1 frame 720x1280 Q=200, CPU oracle operators", "mask_bits_compared": tot,
                  "mask_bits_flipped": flips,
                  "pred_masks_rel_diff": ((pm0 - pm1).abs().max() / pm0.abs().max()).item(),
                  "pred_logits_rel_diff": ((o0["pred_logits"] - o1["pred_logits"]).abs().max() / o0["pred_logits"].abs().max()).item()}))
