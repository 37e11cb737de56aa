"""UniVS_Prompt meta-architecture, inference side (univs/univs_prompt.py:66-490).

Registered under the reference's name so `MODEL.META_ARCHITECTURE: UniVS_Prompt` resolves here.  It owns
`backbone` and `sem_seg_head` (the two calls every task head makes, e.g. inference_video_vis_fast.py:232-236) and the
non-persistent `pixel_mean` / `pixel_std` buffers (:169-170).  `forward(batched_inputs)` keeps the input contract
(list of one dict with "image": list of T [3,H,W] tensors, "height", "width", "task", "dataset_name", ...) and runs
ONE clip through normalise -> pad to a multiple of 32 (ImageList.from_tensors semantics) -> backbone -> sem_seg_head.
Without task heads `forward` returns the per-clip decoder output dict.  With task heads (`attach_task_heads`, or a cfg
that carries MODEL.UniVS.TEST / MODEL.BoxVIS.TEST) `forward` = `forward_inference`: the reference's dispatch over whole
videos (univs_prompt.py:416-452) to the sliding-window heads of `univs_b200/inference/`.

Frame sharding (new capability): with a process group of n ranks, rank r runs backbone + pixel decoder
on frames {f : f mod n == r}; one all-gather reassembles the three multi-scale maps and mask_features; the decoder
then runs on every rank."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import nn_ops, ops, switches
from . import precision  # noqa: F401  (applies the documented default policy at import)
from .registry import META_ARCH_REGISTRY, build_backbone, build_sem_seg_head
from .sharding import FrameSharder, TokenExchange


@META_ARCH_REGISTRY.register()
class UniVS_Prompt(nn.Module):
    def __init__(self, cfg=None, *, backbone=None, sem_seg_head=None, pixel_mean=None, pixel_std=None,
                 size_divisibility=32, num_frames=5, process_group=None, shard_decoder=None, frame_streams=None):
        super().__init__()
        if cfg is not None:
            backbone = build_backbone(cfg)
            sem_seg_head = build_sem_seg_head(cfg, backbone.output_shape())
            pixel_mean, pixel_std = cfg.MODEL.PIXEL_MEAN, cfg.MODEL.PIXEL_STD
            size_divisibility = cfg.MODEL.MASK_FORMER.SIZE_DIVISIBILITY
            num_frames = cfg.INPUT.SAMPLING_FRAME_NUM
        self.backbone = backbone
        self.sem_seg_head = sem_seg_head
        if size_divisibility < 0:
            size_divisibility = self.backbone.size_divisibility
        self.size_divisibility = size_divisibility
        self.num_frames = num_frames
        self.register_buffer("pixel_mean", torch.tensor(pixel_mean, dtype=torch.float32).view(-1, 1, 1), False)
        self.register_buffer("pixel_std", torch.tensor(pixel_std, dtype=torch.float32).view(-1, 1, 1), False)
        self.sharder = FrameSharder(process_group)
        # frame-sharded decoder with per-layer token exchange instead of the feature all-gather (opt-in)
        self.shard_decoder = (switches.get("SHARD_DECODER") == 1) if shard_decoder is None else shard_decoder
        # single GPU: backbone + pixel decoder of g frame groups on g CUDA streams, so that the memory-bound passes of one
        # group can run under the tensor-core GEMMs of another (frames are independent up to the decoder); opt-in
        self.frame_streams = switches.get("FRAME_STREAMS") if frame_streams is None else int(frame_streams)
        self._streams = []
        self._cfg = cfg
        self.task_heads = None                 # see attach_task_heads
        self.eval()

    @property
    def device(self):
        return self.pixel_mean.device

    # ---- pre-processing: (x - mean) / std, right/bottom zero pad to a multiple of size_divisibility
    def preprocess(self, frames):
        """frames: [T,3,H,W] uint8/float (RGB) or a list of [3,H,W].  Returns ([T,3,Hp,Wp] float32, (H, W))."""
        if isinstance(frames, (list, tuple)):
            frames = torch.stack([f.to(self.device, non_blocking=True) for f in frames])
        else:
            frames = frames.to(self.device, non_blocking=True)
        x = (frames.float() - self.pixel_mean) / self.pixel_std
        H, W = x.shape[-2:]
        d = self.size_divisibility
        Hp, Wp = (H + d - 1) // d * d, (W + d - 1) // d * d
        if (Hp, Wp) != (H, W):
            x = F.pad(x, (0, Wp - W, 0, Hp - H), value=0.0)
        return x, (H, W)

    def _padded_size(self, H, W):
        d = self.size_divisibility
        return (H + d - 1) // d * d, (W + d - 1) // d * d

    @torch.no_grad()
    def backbone_from_frames(self, frames):
        """backbone(preprocess(frames)) with the ingest fused into the patch embedding (nn_ops.fused_glue() path)."""
        if isinstance(frames, (list, tuple)):
            frames = torch.stack([f.to(self.device, non_blocking=True) for f in frames])
        else:
            frames = frames.to(self.device, non_blocking=True)
        if frames.dtype not in (torch.uint8, torch.float32):
            frames = frames.float()
        H, W = frames.shape[-2:]
        # host copies of the normalisation constants, read back ONCE per buffer version: a device->host read in every step
        # cannot be captured in a CUDA graph (round 2: this silently sent the fused-glue step down the eager path)
        key = (self.pixel_mean.data_ptr(), self.pixel_mean._version, self.pixel_std.data_ptr(), self.pixel_std._version)
        cached = self.__dict__.get("_mean_std_host")
        if cached is None or cached[0] != key:
            cached = (key, self.pixel_mean.flatten().tolist(), self.pixel_std.flatten().tolist())
            self.__dict__["_mean_std_host"] = cached
        return self.backbone.forward_frames(frames.contiguous(), cached[1], cached[2], self._padded_size(H, W))

    @torch.no_grad()
    def clip_forward(self, frames, targets):
        """One clip through the hot path.  frames [T,3,H,W] (any device / dtype), targets list[dict] (mutated in
        place by the prompt sampler, as in the reference)."""
        fused = nn_ops.fused_glue() and hasattr(self.backbone, "forward_frames")
        if fused:       # ingest (cast, normalise, pad, patchify) happens inside the patch embedding
            if isinstance(frames, (list, tuple)):
                frames = torch.stack([f.to(self.device, non_blocking=True) for f in frames])
            x = None
            T, image_size = frames.shape[0], tuple(frames.shape[-2:])
            padded = self._padded_size(*image_size)
        else:
            x, image_size = self.preprocess(frames)
            T, padded = x.shape[0], tuple(x.shape[-2:])
        targets[0].setdefault("inter_image_size", padded)
        targets[0].setdefault("image_size", image_size)
        targets[0].setdefault("num_frames", T)
        if "frame_indices" in targets[0]:
            targets[0]["frame_indices"] = targets[0]["frame_indices"].to(self.device)
        if self.sharder.world_size == 1:
            if self.frame_streams > 1 and T > 1:
                return self._grouped_forward(frames if fused else x, fused, T, targets)
            features = self.backbone_from_frames(frames) if fused else self.backbone(x)
            return self.sem_seg_head(features, targets=targets)
        pd = self.sem_seg_head.pixel_decoder
        dec = self.sem_seg_head.predictor
        if self.shard_decoder and dec.supports_exchange(targets):
            # frames stay on their rank through the decoder; only query tokens are exchanged (sharding.TokenExchange)
            # one exchange object per clip length: its device-resident frame-index tensors are created once (a host->device
            # copy inside the step would break CUDA-graph capture)
            cache = self.__dict__.setdefault("_exchange_cache", {})
            ex = cache.get(T)
            if ex is None:
                ex = cache[T] = TokenExchange(self.sharder, T)
            whole = frames if fused else x
            local = whole.index_select(0, ex.frames_tensor(whole.device))
            features = self.backbone_from_frames(local) if fused else self.backbone(local)
            mask_features, mf_bfe, _enc, multi_scale = pd.forward_features(features)
            return dec(multi_scale, mask_features, mf_bfe, None, targets, exchange=ex)
        # frame-sharded: local frames -> backbone -> pixel decoder -> all-gather -> decoder
        local = self.sharder.local_frames(frames if fused else x)
        if local.shape[0] > 0:
            features = self.backbone_from_frames(local) if fused else self.backbone(local)
            mask_features, mf_bfe, _enc, multi_scale = pd.forward_features(features)
            parts = [t.permute(0, 2, 3, 1) for t in [mask_features] + list(multi_scale)]
        else:   # more ranks than frames: this rank owns nothing and only takes part in the exchange
            Hp, Wp = padded
            parts = [torch.zeros((0, Hp // s, Wp // s, c), device=self.device) for s, c in
                     [(4, pd.mask_dim), (32, pd.conv_dim), (16, pd.conv_dim), (8, pd.conv_dim)]]
        # exchange in storage order (channel-last), hand NCHW views back to the decoder
        gathered = self.sharder.all_gather_frames(parts, T)
        gathered = [t.permute(0, 3, 1, 2) for t in gathered]
        mask_features, multi_scale = gathered[0], gathered[1:]
        return self.sem_seg_head.predictor(multi_scale, mask_features, mask_features, None, targets)

    def _grouped_forward(self, whole, fused, T, targets):
        """Backbone + pixel decoder per contiguous frame group, each group on its own CUDA stream (sequentially on CPU
        tensors); the per-group outputs are concatenated in frame order and the decoder runs once on the current stream.
        Frame independence of this part of the path is the property frame sharding relies on (sharding.py)."""
        n = min(self.frame_streams, T)
        bounds = [(T * g) // n for g in range(n + 1)]
        pd = self.sem_seg_head.pixel_decoder
        on_gpu = self.device.type == "cuda"
        outs = []
        if on_gpu:
            cur = torch.cuda.current_stream()
            while len(self._streams) < n:
                self._streams.append(torch.cuda.Stream(device=self.device))
        for g in range(n):
            part = whole[bounds[g]:bounds[g + 1]]

            def run():
                feats = self.backbone_from_frames(part) if fused else self.backbone(part)
                mask_features, mf_bfe, _enc, multi_scale = pd.forward_features(feats)
                return mask_features, list(multi_scale)

            ops.scratch_slot = g                 # persistent scratch buffers are per group (groups run concurrently)
            try:
                if on_gpu:
                    st = self._streams[g]
                    st.wait_stream(cur)
                    with torch.cuda.stream(st):
                        outs.append(run())
                else:
                    outs.append(run())
            finally:
                ops.scratch_slot = 0
        if on_gpu:
            for g in range(n):
                cur.wait_stream(self._streams[g])
        cat = lambda ts: torch.cat([t.permute(0, 2, 3, 1) for t in ts], 0).permute(0, 3, 1, 2)   # keep channel-last storage
        mask_features = cat([o[0] for o in outs])
        multi_scale = [cat([o[1][l] for o in outs]) for l in range(len(outs[0][1]))]
        # mask_features_bfe_conv is not read by the inference decoder (only viewed, ..._univs.py:313): as in the sharded path
        return self.sem_seg_head.predictor(multi_scale, mask_features, mask_features, None, targets)

    # ---- whole-video inference: the reference's dispatch to the task heads (univs_prompt.py:416-452)
    def attach_task_heads(self, cfg=None, *, thing_ids=(), thing_contiguous_ids=(), metadata=None,
                          video_unified_inference_enable=None,
                          tracker_type=None, custom_videos_enable=None, custom_videos_text=None, **head_kwargs):
        """Builds the task heads this build has (VIS with the MinVIS tracker, VOS / RefVOS, VPS, unified entity head, image
        head) from `cfg` (default: the cfg the model was built from) and makes `forward` dispatch to them.
        `thing_ids` (1-based dataset ids) / `thing_contiguous_ids` (0-based class indices) / `metadata` stand in for detectron2's MetadataCatalog entry of the test dataset; `head_kwargs` are
        passed to every head that accepts them (e.g. reuse_features=False)."""
        import inspect
        from .inference import (InferenceImageGenericSeg, InferenceVideoEntity, InferenceVideoSemanticExtraction,
                                InferenceVideoVISFast, InferenceVideoVOS, InferenceVideoVPS)
        cfg = self._cfg if cfg is None else cfg
        uv = cfg.MODEL.UniVS.TEST if cfg is not None else {}
        bv = cfg.MODEL.BoxVIS.TEST if cfg is not None else {}
        pick = lambda given, node, key, default: given if given is not None else (node.get(key, default) if node else default)

        def make(cls, **extra):
            accepted = set(inspect.signature(cls.__init__).parameters)
            return cls(cfg, **{k: v for k, v in {**head_kwargs, **extra}.items() if k in accepted})

        self.task_heads = {
            "vis_fast": make(InferenceVideoVISFast),
            "vos": make(InferenceVideoVOS, metadata=metadata),
            "vps": make(InferenceVideoVPS, thing_ids=thing_ids),
            "entity": make(InferenceVideoEntity, thing_ids=thing_ids),
            "image": make(InferenceImageGenericSeg, thing_contiguous_ids=thing_contiguous_ids),
            "semantic_extraction": make(InferenceVideoSemanticExtraction),
            "unified": bool(pick(video_unified_inference_enable, uv, "VIDEO_UNIFIED_INFERENCE_ENABLE", False)),
            "custom_videos": bool(pick(custom_videos_enable, uv, "CUSTOM_VIDEOS_ENABLE", False)),
            "custom_videos_text": list(pick(custom_videos_text, uv, "CUSTOM_VIDEOS_TEXT", [])),
            "tracker_type": pick(tracker_type, bv, "TRACKER_TYPE", "minvis"),
        }
        return self

    @torch.no_grad()
    def forward_inference(self, batched_inputs):
        h = self.task_heads
        if h is None:
            raise RuntimeError("forward_inference needs task heads: call attach_task_heads() first")
        if getattr(self.sem_seg_head.predictor, "semantic_extraction_enable", False):
            return h["semantic_extraction"].eval(self, batched_inputs)
        if "dataset_name" not in batched_inputs[0]:
            # demo form (demo/predictor.py:106-118: {"image", "height", "width"} only; the reference's dispatch raises a
            # KeyError on it): category-specified detection over the default video vocabulary
            batched_inputs = [dict(batched_inputs[0], dataset_name=h.get("demo_dataset", "ytvis21"),
                                   task=batched_inputs[0].get("task", "detection"))]
            batched_inputs[0].setdefault("video_len", len(batched_inputs[0]["image"]))
        name = batched_inputs[0]["dataset_name"]
        if name.startswith("coco") or name.startswith("ade20k"):
            return h["image"].eval(self, batched_inputs)                # image vocabularies
        if len(h["custom_videos_text"]) and batched_inputs[0].get("task") not in ("grounding", "sot"):
            # custom-text route (univs_prompt.py:431-433 + prepare_targets.py:66-71): the video becomes a grounding video whose
            # expressions are the configured texts.  The reference encodes them with its CLIP text tower at this point; that
            # tower is outside this build, so the caller supplies the features.
            texts = list(h["custom_videos_text"][0])
            missing = [k for k in ("exp_word_feats", "exp_sentence_feats", "exp_word_len") if k not in batched_inputs[0]]
            if missing:
                raise NotImplementedError(
                    "CUSTOM_VIDEOS_TEXT: the expressions must come with their CLIP text features "
                    f"({', '.join(missing)} missing from the input dict): the CLIP text tower is not part of this build")
            batched_inputs = [dict(batched_inputs[0], task="grounding", expressions=texts, exp_obj_ids=list(range(len(texts))))]
        if batched_inputs[0].get("task") in ("grounding", "sot"):
            return h["vos"].eval(self, batched_inputs)                  # prompt-specified tasks
        if h["unified"] or h["custom_videos"]:                          # category-specified tasks, unified entity inference
            if name.startswith(("ytvis", "ovis", "vipseg", "vspw")) or h["custom_videos"]:
                return h["entity"].eval(self, batched_inputs)
            raise ValueError(f"Not support to eval the dataset {name} yet")
        if name.startswith(("ytvis", "ovis")):
            if h["tracker_type"] == "mdqe":
                raise NotImplementedError("the clip-level MDQE tracker head (inference_video_vis.py) is not part of this build")
            return h["vis_fast"].eval(self, batched_inputs)
        if name.startswith(("vipseg", "vpsw")):
            return h["vps"].eval(self, batched_inputs)
        raise ValueError(f"Not support to eval the dataset {name} yet")

    def forward(self, batched_inputs):
        if self.training:
            raise NotImplementedError("training is out of scope for the B200 hot-path build")
        if self.task_heads is not None:
            return self.forward_inference(batched_inputs)
        assert len(batched_inputs) == 1, "inference processes one video at a time (univs_prompt.py:421)"
        inp = batched_inputs[0]
        frames = inp["image"]
        T = len(frames)
        task = inp.get("task", "detection")
        prompt_type = "text" if task == "grounding" else ("visual" if task in ("detection", "sot") else "visual")
        tg = {"task": task, "dataset_name": inp.get("dataset_name", "ytvis21"), "prompt_type": inp.get("prompt_type", prompt_type),
              "frame_indices": inp.get("frame_indices", torch.arange(T)), "num_frames": T,
              "video_len": inp.get("video_len", T), "file_names": inp.get("file_names", [""] * T)}
        for k in ("masks", "boxes", "ids", "first_appear_frame_idxs", "first_frame_idx", "exp_word_feats",
                  "exp_sentence_feats", "exp_word_len"):
            if k in inp:
                tg[k] = inp[k]
        out = self.clip_forward(frames, [tg])
        out.pop("aux_outputs", None)
        return out
