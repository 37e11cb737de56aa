"""Pins oracle/ops_ref.py (the CPU restatement) against the reference's OWN source files,
executed by path through oracle/ref_shim.py.  Runs only where /root/reference exists."""
import pytest
import torch

from oracle import ops_ref, ref_shim

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    return ref_shim.load()


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)


def test_msda_reference_test_shapes(ref):
    """Replays the shapes/seed/inputs of the reference's only hot-path test, ops/test.py:24-63."""
    torch.manual_seed(3)
    N, M, D = 1, 2, 2
    Lq, L, P = 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    lsi = torch.cat((shapes.new_zeros((1,)), shapes.prod(1).cumsum(0)[:-1]))
    S = int(shapes.prod(1).sum())
    for dt, tol in ((torch.float64, 1e-12), (torch.float32, 1e-5)):
        value = torch.rand(N, S, M, D, dtype=dt) * 0.01
        loc = torch.rand(N, Lq, M, L, P, 2, dtype=dt)
        w = torch.rand(N, Lq, M, L, P, dtype=dt) + 1e-5
        w = w / w.sum(-1, keepdim=True).sum(-2, keepdim=True)
        want = ref.ms_deform_attn_core_pytorch(value, shapes.tolist(), loc, w)
        got = ops_ref.ms_deform_attn(value, shapes.tolist(), lsi.tolist(), loc, w)
        assert _rel(got, want) < tol


@pytest.mark.parametrize("shapes", [[(5, 7), (10, 14), (20, 27)], [(2, 3)], [(1, 1), (9, 4)]])
def test_msda_core_random(ref, shapes):
    torch.manual_seed(0)
    N, M, D, P = 2, 8, 32, 4
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    lsi = [0]
    for h, w in shapes[:-1]:
        lsi.append(lsi[-1] + h * w)
    value = torch.randn(N, S, M, D)
    loc = torch.rand(N, S, M, L, P, 2) * 1.4 - 0.2          # includes out-of-range samples
    w = torch.softmax(torch.randn(N, S, M, L * P), -1).view(N, S, M, L, P)
    want = ref.ms_deform_attn_core_pytorch(value, shapes, loc, w)
    got = ops_ref.ms_deform_attn(value, shapes, lsi, loc, w)
    assert _rel(got, want) < 1e-5


def test_msda_fused_vs_module(ref):
    torch.manual_seed(1)
    shapes = [(3, 5), (6, 10), (12, 20)]
    S = sum(h * w for h, w in shapes)
    mod = ref.MSDeformAttn(256, 3, 8, 4)
    with torch.no_grad():
        mod.sampling_offsets.weight.normal_(std=0.05)
        mod.attention_weights.weight.normal_(std=0.05)
    src = torch.randn(2, S, 256)
    pos = torch.randn(2, S, 256)
    sp = torch.as_tensor(shapes, dtype=torch.long)
    lsi = torch.cat((sp.new_zeros((1,)), sp.prod(1).cumsum(0)[:-1]))
    enc = ref.pix.MSDeformAttnTransformerEncoder
    refpts = enc.get_reference_points(sp, torch.ones(2, 3, 2), "cpu")
    with torch.no_grad():
        want = mod(src + pos, refpts, src, sp, lsi)
        value = mod.value_proj(src).view(2, S, 8, 32)
        q = src + pos
        ol = torch.cat([mod.sampling_offsets(q), mod.attention_weights(q)], -1)
        got = mod.output_proj(ops_ref.ms_deform_attn_fused(value, shapes, lsi.tolist(), ol, 8, 3, 4))
    assert _rel(got, want) < 1e-5


@pytest.mark.parametrize("H,W,ws,nH", [(8, 12, 4, 2), (7, 10, 4, 3), (13, 9, 7, 1), (24, 27, 12, 2)])
def test_swin_window_attention_vs_basic_layer(ref, H, W, ws, nH):
    torch.manual_seed(2)
    C = 32 * nH
    layer = ref.swin.BasicLayer(dim=C, depth=2, num_heads=nH, window_size=ws, drop_path=0.0)
    layer.eval()
    with torch.no_grad():
        for blk in layer.blocks:
            blk.attn.relative_position_bias_table.normal_(std=0.5)
            blk.attn.qkv.bias.normal_(std=0.3)
    x = torch.randn(2, H * W, C)
    with torch.no_grad():
        want = layer(x, H, W)[0]
        y = x
        for i, blk in enumerate(layer.blocks):
            qkv = torch.nn.functional.linear(blk.norm1(y), blk.attn.qkv.weight).view(2, H, W, 3 * C)   # bias-free
            a = ops_ref.swin_window_attention(qkv, blk.attn.qkv.bias, blk.attn.relative_position_bias_table,
                                              nH, ws, 0 if i % 2 == 0 else ws // 2)
            y = y + blk.attn.proj(a.view(2, H * W, C))
            y = y + blk.mlp(blk.norm2(y))
    assert _rel(y, want) < 1e-5


def test_mha_core_vs_cross_attention_layer(ref):
    torch.manual_seed(3)
    Q, S, B, C, h = 11, 37, 3, 256, 8
    layer = ref.CrossAttentionLayer(C, h)
    layer.eval()
    tgt, qpos = torch.randn(Q, B, C), torch.randn(Q, B, C)
    mem, pos = torch.randn(S, B, C), torch.randn(S, B, C)
    mask = torch.rand(B, Q, S) < 0.6
    mask[1, 4] = True     # a fully blocked row -> un-blocked by the caller (..._univs.py:390)
    m_ref = mask[:, None].repeat(1, h, 1, 1).flatten(0, 1).clone()
    m_ref[torch.where(m_ref.sum(-1) == m_ref.shape[-1])] = False
    with torch.no_grad():
        want = layer(tgt, mem, memory_mask=m_ref, pos=pos, query_pos=qpos)
        mh = layer.multihead_attn
        Wq, Wk, Wv = mh.in_proj_weight.chunk(3)
        bq, bk, bv = mh.in_proj_bias.chunk(3)
        q = ((tgt + qpos) @ Wq.T + bq).transpose(0, 1)
        k = ((mem + pos) @ Wk.T + bk).transpose(0, 1)
        v = (mem @ Wv.T + bv).transpose(0, 1)
        a = ops_ref.mha_core(q, k, v, h, mask.to(torch.uint8), unmask_full_rows=True)
        got = layer.norm(tgt + mh.out_proj(a).transpose(0, 1))
    assert _rel(got, want) < 1e-5


def test_attn_mask_from_logits_is_interpolate_threshold(ref):
    torch.manual_seed(4)
    Q, T, H, W = 5, 2, 16, 24
    logits = torch.randn(Q, T, H * W)
    for tgt in ((8, 12), (4, 6), (2, 3)):
        want = torch.nn.functional.interpolate(logits.view(Q, T, H, W), size=tgt, mode="bilinear",
                                               align_corners=False)
        want = (want.permute(1, 0, 2, 3).flatten(2).sigmoid() < 0.5)
        got = ops_ref.attn_mask_from_logits(logits, (H, W), tgt).bool()
        assert torch.equal(got, want)


def test_position_encodings(ref):
    pe2 = ref.pe2d.PositionEmbeddingSine(128, normalize=True)
    x = torch.zeros(1, 4, 9, 13)
    assert _rel(ops_ref.pos2d_sine(9, 13), pe2(x)[0]) < 1e-6
    pe3 = ref.pe3d.PositionEmbeddingSine3DArbitraryT(128, normalize=True)
    fi = torch.tensor([[3, 4, 9]])
    want = pe3(torch.zeros(1, 3, 4, 9, 13), fi)[0]
    assert _rel(ops_ref.pos3d_sine_arbitrary_t(fi[0], 9, 13), want) < 1e-6


def test_proca_core_vs_reference_layer(ref):
    torch.manual_seed(5)
    P, T, L, C, h = 4, 3, 6, 256, 8
    layer = ref.CrossAttentionLayer(C, h)
    layer.eval()
    tok, qe = torch.randn(P, T, C), torch.randn(P, T, C)
    mem, mpe = torch.randn(P, L, T, C), torch.randn(P, L, T, C)
    with torch.no_grad():
        # restated call pattern of ..._univs.py:474-492
        dense = torch.cat([tok.unsqueeze(1), mem], 1).transpose(0, 1).flatten(1, 2)
        dpos = torch.cat([qe.unsqueeze(1), mpe], 1).transpose(0, 1).flatten(1, 2)
        want = layer(tok.flatten(0, 1)[None], dense, pos=dpos, query_pos=qe.flatten(0, 1)[None])[0].view(P, T, C)
        mh = layer.multihead_attn
        Wq, Wk, Wv = mh.in_proj_weight.chunk(3)
        bq, bk, bv = mh.in_proj_bias.chunk(3)
        q = (tok + qe) @ Wq.T + bq
        ks = (tok + qe) @ Wk.T + bk
        vs = tok @ Wv.T + bv
        km = ((mem + mpe) @ Wk.T + bk).permute(0, 2, 1, 3)
        vm = (mem @ Wv.T + bv).permute(0, 2, 1, 3)
        a = ops_ref.proca_core(q, ks, vs, km, vm, h)
        got = layer.norm(tok + mh.out_proj(a))
    assert _rel(got, want) < 1e-5
