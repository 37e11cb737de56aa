"""Mask visual prompts -> dense prompt tokens + the per-video prompt memory pool (inference).

Restates, on token-major device tensors, the inference branches of
univs/modeling/prompt_encoder/prompt_encoder.py: `VisualPromptSampler.process_per_video_inference` (:844-960),
`process_per_video_inference_prev_frame` (:963-1057), `zero_pad_prompt` (:1059-1071) and
`VisualPromptEncoder.get_mask_prompt` (:167-263), `select_points_from_box_mask` (:361-442, inference branch),
`get_dense_features` (:444-497), plus `convert_box_to_mask` / `convert_mask_to_box` (univs/utils/comm.py:6-84).

Random point picking uses the CPU generator in the reference's call order (`torch.randperm(n)` without a device
argument, :420-425, :481), so a seeded run reproduces the reference's choices.  The pool tensors live in the caller's
`targets[0]`: `prompt_feats`, `prompt_pe` [P,R,n_frames,C], `prompt_attn_masks` [n_frames,1,P,hw], `prompt_obj_ids`.
Only `prompt_type == "masks"` (what every inference caller passes, e.g. inference_video_vos.py) is built.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import nn_ops
from . import position


def mask_to_box(masks: torch.Tensor) -> torch.Tensor:
    """XYXY box (pixel indices) around each bool mask [..., H, W]; [0,0,0,0] for empty (comm.py:41-84)."""
    if masks.numel() == 0:
        return torch.zeros(*masks.shape[:-2], 4, device=masks.device)
    h, w = masks.shape[-2:]
    m = masks.reshape(-1, h, w)
    rows, cols = m.any(-1), m.any(-2)
    ar_h = torch.arange(h, device=m.device)
    ar_w = torch.arange(w, device=m.device)
    bottom = (rows * ar_h).max(-1)[0]
    top = (rows * ar_h + h * (~rows)).min(-1)[0]
    right = (cols * ar_w).max(-1)[0]
    left = (cols * ar_w + w * (~cols)).min(-1)[0]
    empty = (right < left) | (bottom < top)
    out = torch.stack([left, top, right, bottom], -1) * (~empty).unsqueeze(-1)
    return out.reshape(*masks.shape[:-2], 4)


def box_to_mask(boxes: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """normalised XYXY boxes [Q,4] -> bool [Q,h,w]; cell (y,x) is inside iff floor(x1*w) < x <= ceil(x2*w) (comm.py:6-39)."""
    b = boxes * torch.as_tensor([w, h, w, h], dtype=boxes.dtype, device=boxes.device)
    x1, y1, x2, y2 = b[:, 0].floor(), b[:, 1].floor(), b[:, 2].ceil(), b[:, 3].ceil()
    gy = torch.arange(h, device=boxes.device).view(1, h, 1)
    gx = torch.arange(w, device=boxes.device).view(1, 1, w)
    return (gx > x1[:, None, None]) & (gx <= x2[:, None, None]) & (gy > y1[:, None, None]) & (gy <= y2[:, None, None])


_coords_cache = {}


def _pixel_centres(h, w, dev):
    """normalised (x, y) pixel centres [h*w, 2], cached per size and device"""
    key = (h, w, str(dev))
    c = _coords_cache.get(key)
    if c is None:
        ys, xs = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
        c = ((torch.stack([xs, ys], -1) + 0.5) / torch.as_tensor([w, h]).view(1, 1, -1)).flatten(0, 1).to(dev)
        _coords_cache[key] = c
    return c


def _kth_true(sel, ranks):
    """index of the ranks[i, j]-th (0-based, row-major) True of row i of `sel` [Q, n] -- what `nonzero(sel[i])[rank]` picks"""
    cs = sel.to(torch.int32).cumsum(1)
    return torch.searchsorted(cs, (ranks + 1).to(torch.int32), right=False).clamp(max=sel.shape[1] - 1)


def _point_candidates(masks, boxes, mask_thresh=0.75):
    """select_points_from_box_mask, inference branch (:361-442), device part: per instance the candidate pixels -- the
    centre quarter of the box inside the mask, or (none there) the pixels at >= min(0.95, max) of the mask."""
    Q, h, w = masks.shape
    mf = masks.float().flatten(1)
    coords = _pixel_centres(h, w, masks.device)
    cxcy = 0.5 * (boxes[:, :2] + boxes[:, 2:])
    wh = boxes[:, 2:] - boxes[:, :2]
    mmax = mf.max(1)[0]
    binary = mf >= mmax.clamp(max=mask_thresh).reshape(-1, 1)
    in_ctr = ((coords[None] - cxcy[:, None]).abs() < 0.25 * wh[:, None]).all(-1) & binary
    fallback = mf >= mmax.clamp(max=0.95).reshape(-1, 1)
    sel = torch.where(in_ctr.any(1, keepdim=True), in_ctr, fallback)
    return sel, coords


def _mask_prompt(sampler, feat, pe, masks, boxes, key_fid, key_fid_original, T, h, w):
    """get_mask_prompt (:167-263) for one key frame.  feat/pe: [hw, C] tokens of the 1/8 level.

    The reference walks the instances in Python (select_points_from_box_mask :361-442, get_dense_features :444-497): a
    `.any()` / `.max()` / `nonzero` host synchronisation and a handful of small device ops per instance.  Here the device
    side is vectorised over the instances and ONE device->host read per key frame brings back the candidate counts; the
    random picks stay `torch.randperm(n)` on the CPU generator, per instance and in the reference's order (points first,
    then dense features), so a seeded run reproduces the reference's choices; the picks go back as one index tensor."""
    dev = feat.device
    R, s = sampler.num_dense_points, sampler.img_feats_scale
    Q, hm, wm = masks.shape
    if (h * s, w * s) != (hm, wm):
        raise AssertionError(f"Input images must have same size with masks: {(hm, wm), (h * s, w * s)}")
    valid = masks.gt(0.5).flatten(1).sum(-1) > 0
    sel, coords = _point_candidates(masks, boxes)
    fm = F.interpolate(masks.float().unsqueeze(1), (h, w), mode="nearest").squeeze(1)          # [Q,h,w]
    fm_bin = fm >= fm.max().clamp(max=0.5)
    fmb = fm_bin.flatten(1)
    counts = torch.stack([sel.sum(1), fmb.sum(1)]).cpu()                                       # the one host sync
    n_sel, n_dense = counts[0].tolist(), counts[1].tolist()
    pick = torch.tensor([int(torch.randperm(n)[:1]) for n in n_sel], dtype=torch.long)          # CPU generator (:420-425)
    ranks = torch.zeros((Q, R), dtype=torch.long)
    for i, n in enumerate(n_dense):                                                             # get_dense_features (:470-492)
        if n == 0:
            continue
        ranks[i] = torch.arange(R) % n if n < R else torch.randperm(n)[:R]
    pts = coords[_kth_true(sel, pick.to(dev).view(-1, 1))[:, 0]]                                # [Q,2]
    t_idx = torch.as_tensor(key_fid_original, device=dev).reshape(-1)[:1].repeat(T)
    q_pe = position.sine_3d_points(pts, t_idx, dev, feat.shape[-1] // 2).transpose(0, 1)        # [Q,T,C]
    wgt = (fm * fm_bin).flatten(1)
    with nn_ops.ieee_fp32():
        key_feat = (wgt @ feat) / wgt.sum(-1).clamp(min=0.5)[:, None]                           # [Q,C]
    q_feat = key_feat[:, None].repeat(1, T, 1)
    attn = torch.zeros((T, 1, Q, h * w), dtype=torch.bool, device=dev)
    attn[key_fid, 0] = ~box_to_mask(boxes, h, w).flatten(-2)
    idx = _kth_true(fmb, ranks.to(dev))                                                         # [Q,R] token indices
    empty = (torch.tensor(n_dense) == 0).to(dev).view(-1, 1, 1)
    dense_f = torch.where(empty, q_feat[:, :1].expand(-1, R, -1), feat[idx])                    # no mask pixel: the query itself
    dense_p = torch.where(empty, q_pe[:, :1].expand(-1, R, -1), pe[idx])
    dense_f = dense_f[:, :, None].repeat(1, 1, T, 1)                                            # [Q,R,T,C]
    dense_p = dense_p[:, :, None].repeat(1, 1, T, 1)
    v = valid.view(-1, 1, 1, 1)                                                                 # blank instances: zero prompt,
    dense_p, dense_f = dense_p * v.to(dense_p.dtype), dense_f * v.to(dense_f.dtype)             # nothing blocked (:255-262)
    attn = attn & valid.view(1, 1, -1, 1)
    return dense_p, dense_f, attn


def _zero_pad(sampler, tg):
    """zero_pad_prompt (:1059-1071): the pool grows by clip_stride frames per clip."""
    if "prompt_feats" not in tg:
        return
    cs = sampler.clip_stride
    z = torch.zeros_like(tg["prompt_pe"][:, :, -cs:])
    tg["prompt_pe"] = torch.cat([tg["prompt_pe"], z], 2)
    tg["prompt_feats"] = torch.cat([tg["prompt_feats"], z], 2)
    tg["prompt_attn_masks"] = torch.cat([tg["prompt_attn_masks"], tg["prompt_attn_masks"][-cs:]])
    tg["prompt_attn_masks"][-cs:] = False


def _prev_frame(sampler, tg, feats, pes, h, w):
    """process_per_video_inference_prev_frame (:963-1057): seeds the pool from the frame(s) before the clip when the
    video is entered mid-way."""
    T = feats.shape[0]
    P = tg["masks"].shape[0]
    prev = max(0, int(tg["first_frame_idx"]) - 1)
    fa = tg["first_appear_frame_idxs"]
    appeared = (fa <= prev) & (fa != -1)
    if not ((sampler.num_frames == 1) or ("prompt_feats" not in tg)) or int(appeared.sum()) == 0:
        return
    cs = sampler.clip_stride
    dev = feats.device
    for key_fid in range(cs):
        col = -(T + cs) + key_fid
        boxes = tg["boxes"][:, col].to(dev)[appeared]
        masks = tg["masks"][:, col].to(dev)[appeared]
        orig = tg["frame_indices"][0] - (cs - key_fid)
        pe_d, f_d, am = _mask_prompt(sampler, feats[key_fid], pes[key_fid], masks, boxes, key_fid, orig, T, h, w)
        if "prompt_feats" not in tg:
            _, R, _, C = pe_d.shape
            tg["prompt_pe"] = torch.zeros([P, R, T + cs, C], device=dev)
            tg["prompt_feats"] = torch.zeros([P, R, T + cs, C], device=dev)
            tg["prompt_attn_masks"] = torch.zeros([T + cs, am.shape[1], P, am.shape[-1]], device=dev).bool()
        tg["prompt_pe"][appeared, :, col] = pe_d[:, :, key_fid]
        tg["prompt_feats"][appeared, :, col] = f_d[:, :, key_fid]
        tg["prompt_attn_masks"][col, :, appeared] = am[key_fid]


@torch.no_grad()
def sample_visual_prompts(sampler, src, pos, size_list, tg):
    """process_per_batch_inference (:781-842) + process_per_video_inference (:844-960).
    Returns (prompt_pe_dense, prompt_feats_dense) each [P,R,T,C] (the last T frames of the pool)."""
    li = sampler.prompt_feature_level_index
    feats, pes = src[li], pos[li]                                 # [T,hw,C]
    if pes.shape[0] != feats.shape[0]:
        pes = pes.expand_as(feats)
    h, w = size_list[li]
    T, _, C = feats.shape
    dev = feats.device
    tg["img_emb_per_video"] = feats.transpose(1, 2).reshape(T, C, h, w)
    tg["pos_emb_per_video"] = pes.transpose(1, 2).reshape(T, C, h, w)
    first = int(tg["first_frame_idx"]) == 0
    frame_indices = tg["frame_indices"]
    if not first:
        _zero_pad(sampler, tg)
        _prev_frame(sampler, tg, feats, pes, h, w)
    boxes = tg["boxes"][:, -T:].to(dev)
    masks = tg["masks"][:, -T:].to(dev)
    update = (1 - int(tg["task"] == "grounding")) if first else T - sampler.clip_stride
    for key_fid in range(update):
        pe_d, f_d, am = _mask_prompt(sampler, feats[key_fid], pes[key_fid], masks[:, key_fid], boxes[:, key_fid],
                                     key_fid, frame_indices[key_fid], T, h, w)
        tg["prompt_obj_ids"] = tg["ids"]
        if first:
            tg["prompt_pe"], tg["prompt_feats"], tg["prompt_attn_masks"] = pe_d, f_d, am
        else:
            s_idx = -T + key_fid
            valid = masks[:, key_fid].flatten(1).sum(1) > 0
            tg["prompt_pe"][valid, :, s_idx:] = pe_d[valid, :, key_fid:]
            tg["prompt_feats"][valid, :, s_idx:] = f_d[valid, :, key_fid:]
            tg["prompt_attn_masks"][s_idx:] = am[key_fid:]
    if "prompt_pe" not in tg:
        return None, None
    # NB the reference's "avoid NaN" block (:829-834) multiplies by `isblank` instead of its negation and is a no-op.
    return tg["prompt_pe"][:, :, -T:], tg["prompt_feats"][:, :, -T:]
