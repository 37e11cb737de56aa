"""N>1 path on CPU: world_size-2 gloo process groups.  (a) the ragged frame all-gather reassembles tensors in global
frame order; (b) a frame-sharded clip forward (operators = CPU oracles) equals the single-process forward."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from univs_b200.sharding import FrameSharder, frame_plan


def test_frame_plan():
    assert frame_plan(5, 2) == [[0, 2, 4], [1, 3]]
    assert frame_plan(5, 8) == [[0], [1], [2], [3], [4], [], [], []]
    assert frame_plan(8, 8) == [[i] for i in range(8)]
    assert frame_plan(1, 2) == [[0], []]


def _init(rank, world, port):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)


def _gather_worker(rank, world, port, T, q):
    _init(rank, world, port)
    torch.manual_seed(0)
    full = [torch.randn(T, 3, 4, 5), torch.randn(T, 7)]
    sh = FrameSharder(dist.group.WORLD)
    idx = frame_plan(T, world)[rank]
    local = [t[idx] for t in full]
    out = sh.all_gather_frames(local, T)
    ok = all(torch.equal(a, b) for a, b in zip(out, full))
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [5, 4, 1])
def test_all_gather_frames_gloo(T):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + T + (os.getpid() % 500)
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, T, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(60) for p in procs]
    assert all(ok for _, ok in res), res


def _model_worker(rank, world, port, q, T=3, shard_decoder=False, prompts=False):
    _init(rank, world, port)
    torch.set_num_threads(2)
    from oracle.cpu_backend import oracle_ops
    from tests import model_factory as mf
    from univs_b200.meta_arch import UniVS_Prompt
    from univs_b200.modeling.head import MaskFormerHead
    from univs_b200.registry import ShapeSpec
    Q = 6
    bb, pix, dec = mf.build_product_model(mf.TINY_SWIN, num_queries=Q, num_frames=T, clip_emb=mf.make_clip_emb(),
                                          enc_layers=1, dec_layers=2, text_prompt_to_image_enable=prompts)
    mf.load_keyed((bb, pix, dec))
    shapes = {f"res{i + 2}": ShapeSpec(channels=32 * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    head = MaskFormerHead(shapes, num_classes=133, pixel_decoder=pix, transformer_predictor=dec)
    g = torch.Generator().manual_seed(7)
    frames = torch.rand(T, 3, 60, 90, generator=g) * 255        # not a multiple of 32 -> exercises the padding
    tg = lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual", "frame_indices": torch.arange(T)}]
    if prompts:     # category prompts: one prompt query per class of the dataset, ProCA, lang->vision cross-attention
        tg = lambda: [{"task": "detection", "dataset_name": "bdd_track", "prompt_type": "text", "frame_indices": torch.arange(T) + 2}]
    kw = dict(backbone=bb, sem_seg_head=head, pixel_mean=[123.675, 116.28, 103.53], pixel_std=[58.395, 57.12, 57.375])
    with oracle_ops():
        sharded = UniVS_Prompt(process_group=dist.group.WORLD, shard_decoder=shard_decoder, **kw).clip_forward(frames, tg())
        single = UniVS_Prompt(**kw).clip_forward(frames, tg())
    err = 0.0
    for k in ("pred_masks", "pred_logits", "pred_embds"):
        assert sharded[k].shape == single[k].shape, (k, sharded[k].shape, single[k].shape)
        err = max(err, (sharded[k] - single[k]).abs().max().item() / single[k].abs().max().item())
    q.put((rank, err, tuple(single["pred_masks"].shape)))
    dist.destroy_process_group()


@pytest.mark.parametrize("T", [3, 1])      # T=1 on 2 ranks: rank 1 owns no frame (the 8-GPU / T=5 situation)
def test_frame_sharded_clip_forward_gloo(T):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + T + (os.getpid() % 500)
    procs = [ctx.Process(target=_model_worker, args=(r, 2, port, q, T)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=300) for _ in procs]
    [p.join(60) for p in procs]
    for rank, err, shape in res:
        assert shape == (1, 6, T, 16, 24)
        assert err < 1e-5, (rank, err)


@pytest.mark.parametrize("T,prompts", [(3, False), (1, False), (3, True)])   # T=1: rank 1 shadows frame 0
def test_token_exchange_decoder_gloo(T, prompts):
    """Frame-sharded decoder (per-layer token all-gather instead of the feature all-gather) == single process."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30300 + T + 7 * int(prompts) + (os.getpid() % 500)
    procs = [ctx.Process(target=_model_worker, args=(r, 2, port, q, T, True, prompts)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=300) for _ in procs]
    [p.join(60) for p in procs]
    nq = 6 + (8 if prompts else 0)          # bdd_track: 8 category prompts
    for rank, err, shape in res:
        assert shape == (1, nq, T, 16, 24)
        assert err < 1e-5, (rank, err)
