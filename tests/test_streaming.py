"""Cross-clip reuse: per-frame cached pixel-decoder outputs + per-clip decoder equal the
per-clip recomputation of the whole path (operators = CPU oracles)."""
import torch

from oracle.cpu_backend import oracle_ops
from tests import model_factory as mf
from univs_b200.meta_arch import UniVS_Prompt
from univs_b200.modeling.head import MaskFormerHead
from univs_b200.registry import ShapeSpec
from univs_b200.streaming import ClipStream


def test_clip_stream_equals_per_clip_recompute():
    T, Q, V = 3, 6, 5
    bb, pix, dec = mf.build_product_model(mf.TINY_SWIN, num_queries=Q, num_frames=T, clip_emb=mf.make_clip_emb(),
                                          enc_layers=1, dec_layers=2)
    mf.load_keyed((bb, pix, dec))
    shapes = {f"res{i + 2}": ShapeSpec(channels=32 * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    head = MaskFormerHead(shapes, num_classes=133, pixel_decoder=pix, transformer_predictor=dec)
    model = UniVS_Prompt(backbone=bb, sem_seg_head=head, pixel_mean=[123.675, 116.28, 103.53],
                         pixel_std=[58.395, 57.12, 57.375])
    g = torch.Generator().manual_seed(3)
    video = torch.rand(V, 3, 60, 90, generator=g) * 255
    mk = lambda s: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual",
                     "frame_indices": torch.arange(s, s + T)}]
    with oracle_ops():
        stream = ClipStream(model, T)
        got = {s: o["pred_masks"].clone() for s, o in stream.run(video, mk, stride=1, chunk=2)}
        want = {s: model.clip_forward(video[s:s + T], mk(s))["pred_masks"] for s in range(V - T + 1)}
    assert sorted(got) == [0, 1, 2]
    assert stream.frames_encoded == V                       # every frame through backbone + pixel decoder exactly once
    for s in want:
        assert got[s].shape == want[s].shape
        err = (got[s] - want[s]).abs().max().item() / want[s].abs().max().item()
        assert err < 1e-4, (s, err)


def test_clip_stream_window_larger_than_its_capacity():
    """Heads push NUM_FRAMES_WINDOW_TEST frames at a time; reference configs have W > 2T (T=2 / W=5, T=3 / W=5): frames a
    clip still needs must survive the push (video_vis_fast.py:97)."""
    class _Dec(torch.nn.Module):
        def forward(self, ms, mf, bfe, mask, targets):
            return {"frames": mf[:, 0, 0, 0].clone()}

    class _Pix:
        @staticmethod
        def forward_features(feats):
            x = feats["x"]
            return x, x, x, [x, x, x]

    class _Model:
        sem_seg_head = type("H", (), {"pixel_decoder": _Pix(), "predictor": _Dec()})()
        backbone = staticmethod(lambda x: {"x": x})

    for T, Wn, V in ((2, 5, 11), (3, 5, 9), (2, 7, 8)):
        stream = ClipStream(_Model(), T)                     # default capacity 2T < W
        video = torch.arange(V, dtype=torch.float32).view(V, 1, 1, 1).expand(V, 4, 2, 2).contiguous()
        pushed = 0
        for i in range(V - T + 1):
            while pushed < i + T:
                k = min(Wn, V - pushed)
                stream.push_preprocessed(pushed, video[pushed:pushed + k])
                pushed += k
            out = stream.clip(i, [{}])
            assert out["frames"].tolist() == list(range(i, i + T))
            assert len(stream._cache) <= max(stream.capacity, Wn + T)


def test_frame_groups_equal_whole_clip():
    """UniVS_Prompt(frame_streams=g): backbone + pixel decoder per contiguous frame group (one CUDA stream per group on the
    GPU, sequential on CPU), decoder once -- must equal the ungrouped forward (frames are independent up to the decoder)."""
    from univs_b200 import nn_ops
    T, Q = 5, 6
    bb, pix, dec = mf.build_product_model(mf.TINY_SWIN, num_queries=Q, num_frames=T, clip_emb=mf.make_clip_emb(),
                                          enc_layers=1, dec_layers=2)
    mf.load_keyed((bb, pix, dec))
    shapes = {f"res{i + 2}": ShapeSpec(channels=32 * 2 ** i, stride=4 * 2 ** i) for i in range(4)}
    head = MaskFormerHead(shapes, num_classes=133, pixel_decoder=pix, transformer_predictor=dec)
    kw = dict(backbone=bb, sem_seg_head=head, pixel_mean=[123.675, 116.28, 103.53], pixel_std=[58.395, 57.12, 57.375])
    g = torch.Generator().manual_seed(5)
    frames = (torch.rand(T, 3, 60, 90, generator=g) * 255).round()
    tg = lambda: [{"task": "detection", "dataset_name": "ytvis21", "prompt_type": "visual", "frame_indices": torch.arange(T)}]
    for fused in (False, True):
        nn_ops.set_fused_glue(fused)
        try:
            with oracle_ops("tf32x3" if fused else "fp32"):
                want = UniVS_Prompt(**kw).clip_forward(frames, tg())
                for groups in (2, 3, 8):
                    got = UniVS_Prompt(frame_streams=groups, **kw).clip_forward(frames, tg())
                    for k in ("pred_masks", "pred_logits", "pred_embds"):
                        assert got[k].shape == want[k].shape
                        err = (got[k] - want[k]).abs().max().item() / want[k].abs().max().item()
                        assert err < 1e-4, (fused, groups, k, err)
        finally:
            nn_ops.set_fused_glue(False)


def test_l2_chunked_mlp_equals_the_whole_tensor_schedule():
    """UNIVS_MLP_CHUNK_MB: the Swin MLP in row chunks (fc1 -> GELU -> fc2 per chunk) is the same function of its input"""
    import torch.nn as nn
    from oracle.cpu_backend import oracle_ops
    from univs_b200 import nn_ops
    g = torch.Generator().manual_seed(0)
    fc1, fc2 = nn.Linear(48, 192), nn.Linear(192, 48)
    x = torch.randn(2, 37, 29, 48, generator=g)
    for policy in ("fp32", "tf32x3"):
        with oracle_ops(policy):
            h = nn_ops.prep(x)
            want = nn_ops.mlp(h, fc1, fc2)
            nn_ops.set_mlp_chunk_mb(1)
            try:
                rows = 2 * 37 * 29
                assert nn_ops.mlp_rows_per_chunk(rows, 192) == 256 * ((1 << 20) // (192 * (4 + (8 if policy == "tf32x3" else 4))) // 256) < rows
                got = nn_ops.mlp(h, fc1, fc2)
            finally:
                nn_ops.set_mlp_chunk_mb(0)
        assert got.shape == want.shape == (2, 37, 29, 48)
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(want, fc2(torch.nn.functional.gelu(fc1(x))) - fc2.bias, rtol=1e-4, atol=1e-5)
    assert nn_ops.mlp_rows_per_chunk(1000, 768) == 1000          # switched off: one chunk
