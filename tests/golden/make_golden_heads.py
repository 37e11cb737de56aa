"""Golden results of the REFERENCE's sliding-window task heads (univs/inference/*, executed by path through
oracle/ref_shim.py) on seeded synthetic videos.  Run here (where /root/reference exists):
    python tests/golden/make_golden_heads.py
Weights are key-seeded (tests/model_factory.py); the fixtures hold only the reference's results."""
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from tests import model_factory as mf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
MEAN, STD = [123.675, 116.28, 103.53], [58.395, 57.12, 57.375]

VIS = dict(T=2, Q=12, V=5, H=60, W=90, out=(75, 120), topk=10, model=dict(enc_layers=1, dec_layers=3), seed=5)
SOT = dict(T=3, Q=8, V=6, H=60, W=90, model=dict(enc_layers=1, dec_layers=3, num_dense_points=8, num_prev_frames_memory=4),
           seed=9, rng=21, given={0: [4, 7], 2: [5]})


def video(spec):
    g = torch.Generator().manual_seed(spec["seed"])
    return [(torch.rand(3, spec["H"], spec["W"], generator=g) * 255).round() for _ in range(spec["V"])]


def rects(ids, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    masks, boxes = torch.zeros(len(ids), H, W), torch.zeros(len(ids), 4)
    for j in range(len(ids)):
        bw, bh = int(W * (0.25 + 0.3 * torch.rand(1, generator=g))), int(H * (0.25 + 0.3 * torch.rand(1, generator=g)))
        x0, y0 = int((W - bw) * torch.rand(1, generator=g)), int((H - bh) * torch.rand(1, generator=g))
        masks[j, y0:y0 + bh, x0:x0 + bw] = 1.0
        boxes[j] = torch.tensor([x0, y0, x0 + bw, y0 + bh], dtype=torch.float32)
    return masks, boxes


def sot_annotations(spec, make):
    out = []
    for f in range(spec["V"]):
        ids = spec["given"].get(f, [])
        m, b = rects(ids, spec["H"], spec["W"], 50 + f)
        out.append(make(f, ids, m, b))
    return out


def _reference(spec):
    bb, pix, dec = ref_shim.build_reference_model(mf.TINY_SWIN, num_queries=spec["Q"], num_frames=spec["T"],
                                                  clip_emb=mf.make_clip_emb(), **spec["model"])
    for m in (bb, pix, dec):
        m.load_state_dict(mf.keyed_state_dict(m.state_dict()))
    return bb, pix, dec


def _padded(frames, H, W):
    mean, std = torch.tensor(MEAN).view(3, 1, 1), torch.tensor(STD).view(3, 1, 1)
    xs = torch.stack([(f - mean) / std for f in frames])
    Hp, Wp = (H + 31) // 32 * 32, (W + 31) // 32 * 32
    return torch.nn.functional.pad(xs, (0, Wp - W, 0, Hp - H)), (Hp, Wp)


def _common(T, Q):
    return dict(hidden_dim=256, num_queries=Q, object_mask_threshold=0.05, overlap_threshold=0.8,
                stability_score_thresh=0.0, metadata=None, size_divisibility=32, LSJ_aug_image_size=1024,
                LSJ_aug_enable_test=False, sem_seg_postprocess_before_inference=False, pixel_mean=MEAN, pixel_std=STD,
                num_frames=T, num_classes=133, data_name="val", prompt_as_queries=True, zero_shot_inference=False,
                semantic_on=False, instance_on=True, panoptic_on=False, test_topk_per_image=10, tracker_type="minvis",
                window_inference=False, num_frames_window_test=T, clip_stride=1)


ENTITY = dict(T=2, Q=12, V=7, H=60, W=90, out=(75, 120), model=dict(enc_layers=1, dec_layers=3), seed=21, rng=3,
              head=dict(overlap_threshold=0.3, overlap_threshold_entity=0.2, test_topk_per_image=8, apply_cls_thres=0.45,
                        box_nms_thresh=1.01, clip_stride=1, num_prev_frames_memory=1, temporal_consistency_threshold=0.05,
                        detect_newly_object_threshold=0.05, detect_newly_interval_frames=1))
VPS = dict(T=2, Q=12, V=4, H=60, W=90, out=(75, 120), model=dict(enc_layers=1, dec_layers=3), seed=11,
           things=[c for c in range(1, 125) if c % 3 == 0],
           head=dict(object_mask_threshold=0.0, overlap_threshold=0.3, test_topk_per_image=8))


def main_entity_vps():
    """Fixtures of the unified entity head (VIS sub-task) and the online VPS head."""
    import contextlib
    import inspect
    import io
    import types
    heads = ref_shim.load_inference_heads()
    s = ENTITY
    frames = video(s)
    xs, _ = _padded(frames, s["H"], s["W"])
    kw = dict(_common(s["T"], s["Q"]), is_multi_cls=True, merge_on_cpu=False, num_max_inst_test=50, output_dir="/tmp",
              video_unified_inference_entities="", custom_videos_enable=False, **s["head"])
    accepted = set(inspect.signature(heads.InferenceVideoEntity.__init__).parameters)
    head = heads.InferenceVideoEntity(**{k: v for k, v in kw.items() if k in accepted})
    inputs = [{"image": frames, "height": s["out"][0], "width": s["out"][1], "dataset_name": "ytvis21",
               "task": "detection", "video_len": s["V"], "video_id": 7}]
    tg = [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual", "video_len": s["V"], "sub_task": "vis"}]
    model = heads.RefModel(*_reference(s))
    torch.manual_seed(s["rng"])              # after the model is built: the sampler draws from the global generator
    with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
        res = head.inference_video(model, inputs, heads.ImageList(xs, [(s["H"], s["W"])] * s["V"]), tg)
    res = sorted(res, key=lambda r: (r["category_id"], round(r["score"], 4)))
    torch.save({"category_ids": [r["category_id"] for r in res], "scores": torch.tensor([r["score"] for r in res]),
                "segmentations": [[seg["counts"] for seg in r["segmentations"]] for r in res],
                "ids": tg[0]["ids"], "first_appear": tg[0]["first_appear_frame_idxs"], "logits": tg[0]["logits"],
                "occurrence": tg[0]["occurrence"]}, os.path.join(HERE, "head_entity_vis.pt"))
    print("head_entity_vis", len(res), "entries,", int(tg[0]["ids"].numel()), "entities")

    s = VPS
    frames = video(s)
    xs, _ = _padded(frames, s["H"], s["W"])
    meta = types.SimpleNamespace(thing_dataset_id_to_contiguous_id={c: i for i, c in enumerate(s["things"])})
    kw = dict(_common(s["T"], s["Q"]), overlap_threshold_entity=0.5, is_multi_cls=True, apply_cls_thres=0.05,
              merge_on_cpu=False, num_max_inst_test=50)
    kw.update(s["head"], metadata=meta, data_name="vipseg_val", panoptic_on=True, instance_on=False)
    accepted = set(inspect.signature(heads.InferenceVideoVPS.__init__).parameters)
    head = heads.InferenceVideoVPS(**{k: v for k, v in kw.items() if k in accepted})
    head.change_to_720p = False              # results at the requested output size: keeps the fixture small
    inputs = [{"image": frames, "height": s["out"][0], "width": s["out"][1], "dataset_name": "vipseg", "task": "detection",
               "video_len": s["V"]}]
    with torch.no_grad():
        res = head.inference_video_vps_online(heads.RefModel(*_reference(s)), inputs,
                                              heads.ImageList(xs, [(s["H"], s["W"])] * s["V"]),
                                              [{"task": "detection", "dataset_name": "vipseg", "prompt_type": "visual"}])
    torch.save({"pred_masks": res["pred_masks"].to(torch.int16), "segments_infos": res["segments_infos"],
                "pred_ids": [int(i) for i in res["pred_ids"]], "image_size": res["image_size"]},
               os.path.join(HERE, "head_vps.pt"))
    print("head_vps", len(res["segments_infos"]), "segments", tuple(res["pred_masks"].shape))


def main():
    heads = ref_shim.load_inference_heads()
    # ---- VIS, MinVIS tracker
    s = VIS
    frames = video(s)
    xs, _ = _padded(frames, s["H"], s["W"])
    head = heads.InferenceVideoVISFast(**_common(s["T"], s["Q"]), mdqe_tracker=None, is_multi_cls=True,
                                       apply_cls_thres=0.05, merge_on_cpu=False, num_max_inst_test=50)
    inputs = [{"image": frames, "height": s["out"][0], "width": s["out"][1], "dataset_name": "ytvis21"}]
    with torch.no_grad():
        res = head.inference_video_vis_minvis(
            heads.RefModel(*_reference(s)), inputs, heads.ImageList(xs, [(s["H"], s["W"])] * s["V"]),
            [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual"}])
    order = np.lexsort((res["pred_labels"], res["pred_scores"]))
    torch.save({"scores": torch.tensor([res["pred_scores"][i] for i in order]),
                "labels": torch.tensor([res["pred_labels"][i] for i in order]),
                "masks_packed": torch.from_numpy(np.packbits(np.stack([res["pred_masks"][i].numpy() for i in order]))),
                "masks_shape": (len(order), s["V"], *s["out"])}, os.path.join(HERE, "head_vis_fast.pt"))
    print("head_vis_fast", len(order), "detections")

    # ---- VOS, task sot
    s = SOT
    frames = video(s)
    xs, padded = _padded(frames, s["H"], s["W"])
    names = [f"davis/vid0/{i:05d}.jpg" for i in range(s["V"])]
    with tempfile.TemporaryDirectory() as tmp:
        head = heads.InferenceVideoVOS(**_common(s["T"], s["Q"]), overlap_threshold_entity=0.5, output_dir=tmp,
                                       video_unified_inference_queries="prompt", num_prev_frames_memory=4)
        inst = sot_annotations(s, lambda f, ids, m, b: heads.Instances(
            (s["H"], s["W"]), ori_ids=ids, gt_masks=heads.BitMasks(m), gt_boxes=heads.Boxes(b),
            gt_classes=torch.zeros(len(ids), dtype=torch.long)))
        tg = [{"task": "sot", "dataset_name": "davis", "prompt_type": "visual", "video_len": s["V"], "num_frames": s["T"],
               "inter_image_size": padded, "image_size": (s["H"], s["W"]), "file_names": names, "instances": inst,
               "mask_palette": list(range(256)) * 3}]
        model = heads.RefModel(*_reference(s))
        torch.manual_seed(s["rng"])          # the prompt sampler draws its points from the global CPU generator
        with torch.no_grad():
            head.inference_video_vos(model, None,
                                     heads.ImageList(xs, [(s["H"], s["W"])] * s["V"]), tg, (s["H"], s["W"]), (s["H"], s["W"]))
        from PIL import Image
        d = os.path.join(tmp, "inference/Annotations/vid0")
        maps = {int(n[:5]): torch.from_numpy(np.array(Image.open(os.path.join(d, n)))) for n in sorted(os.listdir(d))}
    torch.save({"id_maps": maps, "boxes": tg[0]["boxes"], "embds": tg[0]["embds"], "ids": tg[0]["ids"]},
               os.path.join(HERE, "head_vos_sot.pt"))
    print("head_vos_sot", sorted(maps), {k: int((v > 0).sum()) for k, v in maps.items()})


if __name__ == "__main__":
    if "entity_vps" in sys.argv[1:]:        # added later in round 1: leaves the first fixtures untouched
        main_entity_vps()
    else:
        main()
