"""Whole-layer / whole-decoder parity of the module mirror against the CPU oracle's restatement of the
reference decoder (oracle/ref_torch.py), same weights (reference state-dict keys), same inputs."""
import copy

import numpy as np
import pytest
import torch

from oracle import ref_torch as R

pytestmark = pytest.mark.gpu


def _setup(name, T, B, Q=None, seed=0, num_layers=2, memory_format='nchw'):
    import sparsebev_b200 as sb
    from sparsebev_b200 import synthetic as S
    cfg = S.layer_cfg(name, T, num_layers=num_layers)
    if Q is not None:
        cfg['num_query'] = Q
    sd = S.make_state_dict(cfg, seed=seed)
    model = sb.SparseBEVTransformer(256, num_frames=T, num_points=cfg['num_points'], num_layers=num_layers,
                                    num_levels=cfg['num_levels'], num_classes=10, code_size=10, pc_range=cfg['pc_range'])
    model.load_state_dict({'decoder.decoder_layer.' + k: v for k, v in sd.items()})
    model = model.cuda().eval()
    feats = S.make_feats(name, T, batch=B, seed=seed + 1, memory_format=memory_format)
    metas = S.make_metas(name, T, batch=B)
    n = int(np.ceil(np.sqrt(cfg['num_query']))) ** 2
    qb = S.init_query_bbox(n, seed=seed + 2)[:cfg['num_query']][None].repeat(B, 1, 1).contiguous()
    qb[..., 8:10] = 0.3 * torch.randn(B, cfg['num_query'], 2, generator=torch.Generator().manual_seed(seed + 3))
    qf = torch.randn(B, cfg['num_query'], 256, generator=torch.Generator().manual_seed(seed + 4))
    return cfg, sd, model, feats, metas, qb, qf


def _rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize('name,T,B', [('tiny', 2, 2), ('tiny', 8, 1), ('tiny5', 3, 1)])
def test_decoder_layer_vs_oracle(name, T, B):
    cfg, sd, model, feats, metas, qb, qf = _setup(name, T, B)
    td = R.time_diff_from_timestamps([m['img_timestamp'] for m in metas])
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas]).astype(np.float32))
    taps = {}
    grouped = R.regroup_feats(feats, channel_last=True)
    want_q, want_cls, want_box = R.decoder_layer(qb, qf, grouped, sd, cfg, td, l2i, op=R.msmv_sampling_kernel_semantics, taps=taps)

    layer = model.decoder.decoder_layer
    metas_gpu = copy.deepcopy(metas)
    model.decoder.prepare_metas(metas_gpu, B, torch.device('cuda'))
    gfeats = model.decoder.prepare_feats([f.cuda() for f in feats])
    got_q, got_cls, got_box = layer(qb.cuda(), qf.cuda(), gfeats, None, metas_gpu)
    assert torch.equal(metas_gpu[0]['time_diff'].cpu(), td)
    assert _rel(got_q, want_q) < 5e-5, 'query_feat rel err %.3e' % _rel(got_q, want_q)
    assert _rel(got_cls, want_cls) < 5e-5, 'cls rel err %.3e' % _rel(got_cls, want_cls)
    assert _rel(got_box, want_box) < 5e-5, 'bbox rel err %.3e' % _rel(got_box, want_box)


@pytest.mark.parametrize('nsplit', [2, 4])
def test_decoder_layer_with_nsplit_cluster_chain(nsplit):
    """Whole layer with the dense chains on the N-split cluster kernel (option dense_nsplit): same parity bar, and
    agreement with the default chain kernel to fp32 round-off of the bf16x3 products."""
    from sparsebev_b200 import _lib
    cfg, sd, model, feats, metas, qb, qf = _setup('tiny', 4, 2, seed=3)
    td = R.time_diff_from_timestamps([m['img_timestamp'] for m in metas])
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas]).astype(np.float32))
    grouped = R.regroup_feats(feats, channel_last=True)
    want_q, want_cls, want_box = R.decoder_layer(qb, qf, grouped, sd, cfg, td, l2i, op=R.msmv_sampling_kernel_semantics)
    layer = model.decoder.decoder_layer
    metas_gpu = copy.deepcopy(metas)
    model.decoder.prepare_metas(metas_gpu, 2, torch.device('cuda'))
    gfeats = model.decoder.prepare_feats([f.cuda() for f in feats])
    default = _lib.get_option('dense_nsplit')
    try:
        _lib.set_option('dense_nsplit', 0)
        base = [t.clone() for t in layer(qb.cuda(), qf.cuda(), gfeats, None, metas_gpu)]
        _lib.set_option('dense_nsplit', nsplit)
        got = layer(qb.cuda(), qf.cuda(), gfeats, None, metas_gpu)
        torch.cuda.synchronize()
    finally:
        _lib.set_option('dense_nsplit', default)
    for g, w, b, what in zip(got, (want_q, want_cls, want_box), base, ('query_feat', 'cls', 'bbox')):
        assert _rel(g, w) < 5e-5, '%s rel err %.3e' % (what, _rel(g, w))
        assert _rel(g, b) < 5e-5, '%s vs default chain kernel: %.3e' % (what, _rel(g, b))


def test_decoder_stages_vs_oracle():
    """Stage-wise: SASA block, sampling block and mixing block each against the oracle taps."""
    T, B = 4, 1
    cfg, sd, model, feats, metas, qb, qf = _setup('tiny', T, B, seed=5)
    td = R.time_diff_from_timestamps([m['img_timestamp'] for m in metas])
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas]).astype(np.float32))
    taps = {}
    grouped = R.regroup_feats(feats, channel_last=True)
    R.decoder_layer(qb, qf, grouped, sd, cfg, td, l2i, op=R.msmv_sampling_kernel_semantics, taps=taps)
    layer = model.decoder.decoder_layer
    metas_gpu = copy.deepcopy(metas)
    model.decoder.prepare_metas(metas_gpu, B, torch.device('cuda'))
    gfeats = model.decoder.prepare_feats([f.cuda() for f in feats])
    q_in = taps['after_sasa'].cuda()
    sampled = layer.sampling(qb.cuda(), q_in, gfeats, metas_gpu)
    assert _rel(sampled, taps['sampled']) < 1e-4, 'sampling block rel err %.3e' % _rel(sampled, taps['sampled'])
    mixed = layer.mixing.forward_fused(taps['sampled'].cuda(), q_in, layer.norm2)
    assert _rel(mixed, taps['mixed']) < 1e-4, 'mixing block rel err %.3e' % _rel(mixed, taps['mixed'])
    layer.mixing.precision = 'bf16'
    fast = layer.mixing.forward_fused(taps['sampled'].cuda(), q_in, layer.norm2)
    layer.mixing.precision = 'bf16x3'
    assert _rel(fast, taps['mixed']) < 3e-2       # single-pass bf16 is a documented fast mode, not the parity mode


def test_full_decoder_and_head_vs_oracle_and_nhwc_zero_copy():
    import sparsebev_b200 as sb
    from sparsebev_b200 import synthetic as S
    T, B, L = 2, 1, 3
    cfg, sd, model, feats, metas, qb, qf = _setup('tiny', T, B, seed=9, num_layers=L)
    td = R.time_diff_from_timestamps([m['img_timestamp'] for m in metas])
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas]).astype(np.float32))
    want_cls, want_box = R.decoder(qb, qf, feats, sd, cfg, td, l2i, channel_last=True, op=R.msmv_sampling_kernel_semantics)
    got_cls, got_box = model(qb.cuda(), qf.cuda(), [f.cuda() for f in feats], None, copy.deepcopy(metas))
    assert got_cls.shape == (L, B, cfg['num_query'], 10) and got_box.shape == (L, B, cfg['num_query'], 10)
    assert _rel(got_cls, want_cls) < 2e-4 and _rel(got_box, want_box) < 2e-4, (_rel(got_cls, want_cls), _rel(got_box, want_box))   # 3 chained layers
    # channels-last features take the zero-copy 'nhwc' path and must agree exactly with the regrouped path
    nhwc = [f.cuda().permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3) for f in feats]
    ptrs = [f.data_ptr() for f in nhwc]
    got2 = model(qb.cuda(), qf.cuda(), nhwc, None, copy.deepcopy(metas))
    assert model.decoder.decoder_layer.sampling.feat_layout == 'nhwc' and [f.data_ptr() for f in nhwc] == ptrs
    assert torch.equal(got2[0], got_cls) and torch.equal(got2[1], got_box)
    # head (eval path)
    head = sb.SparseBEVHead(num_classes=10, in_channels=256, num_query=cfg['num_query'], pc_range=cfg['pc_range'],
                            transformer=dict(type='SparseBEVTransformer', embed_dims=256, num_frames=T, num_points=4,
                                             num_layers=L, num_levels=cfg['num_levels'], pc_range=cfg['pc_range'])).cuda().eval()
    head.transformer.load_state_dict(model.state_dict())
    outs = head([f.cuda() for f in feats], copy.deepcopy(metas))
    want = R.head_forward(head.init_query_bbox.weight.detach().cpu(), head.label_enc.weight.detach().cpu(), feats, sd, cfg,
                          td, l2i, channel_last=True, op=R.msmv_sampling_kernel_semantics)
    assert _rel(outs['all_cls_scores'], want['all_cls_scores']) < 2e-4
    assert _rel(outs['all_bbox_preds'], want['all_bbox_preds']) < 2e-4
    # post-processing on the device tensors: same boxes as decoding the oracle's predictions (top-k ties aside, scores match)
    head.bbox_coder = sb.NMSFreeCoder(pc_range=cfg['pc_range'], post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], max_num=20, num_classes=10)
    dets = head.get_bboxes(outs)
    ref_dets = sb.NMSFreeCoder(pc_range=cfg['pc_range'], post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], max_num=20,
                               num_classes=10).decode({k: v for k, v in want.items() if v is not None})
    assert len(dets) == B
    for (bb, sc, lb), r in zip(dets, ref_dets):
        assert bb.shape[1] == 9 and sc.shape == lb.shape and lb.dtype == torch.int64
        assert sc.shape == r['scores'].shape and torch.allclose(sc.cpu(), r['scores'], rtol=1e-3, atol=1e-4)
        assert torch.equal(lb.cpu()[:3], r['labels'][:3])


# Whole-layer bars: every tensor-core stage is bf16x3 (fp32-grade), smoke measures ~2e-5 on the tiny layer; 5e-5 of the
# output scale leaves 2.5x head-room and is 4x tighter than round 1's 2e-4.
LAYER_BAR = 5e-5


def _fragile_queries(points, l2i, image_h, image_w, delta=1e-4):
    """Queries with a sample point within `delta` of a camera's validity border (u, v in (0,1), depth > eps) in ANY view.
    The first-valid-view pick (sparsebev_sampling.py:102-106) is a discontinuous function of the projected point, and the
    points themselves come out of Linear layers that differ from the oracle's fp32 by ~1e-5 relative (bf16x3 tensor-core
    products): for such a query the two implementations may legitimately sample different cameras.  At 900 queries x 128
    points a handful of queries per layer are in that situation; they are the only rows allowed to miss the parity bar."""
    _, _, cam, _ = R.project_and_select_view(points.reshape(points.shape[0], points.shape[1], points.shape[2], -1, 3), l2i, image_h, image_w, return_all=True)
    u, v, depth = cam[..., 0], cam[..., 1], cam[..., 2]                     # [B,T,N,Q,GP]
    near = (u.abs() < delta) | ((u - 1).abs() < delta) | (v.abs() < delta) | ((v - 1).abs() < delta) | ((depth - 1e-5).abs() < delta)
    near &= (u > -0.5) & (u < 1.5) & (v > -0.5) & (v < 1.5)                 # only borders of a view the point is actually close to
    return near.any(dim=-1).any(dim=1).any(dim=1)                           # [B,Q]


def _rows_within(got, want, bar, fragile, what):
    """Row-wise parity: of the query rows WITHOUT a view-border point, >= 99 % are within `bar` of the output scale and all
    within 10 x bar (the tail is the bilinear taps' sensitivity to the ~1e-6 relative round-off of the sample points, not
    a kernel error: on identical points the gather is exact to 2e-7); rows WITH a view-border point may have picked the
    other camera (at most 2 % of all rows do)."""
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    B, Q = want.shape[:2]
    err = (got - want).reshape(B, Q, -1).abs().amax(-1) / (want.abs().max() + 1e-12)          # [B,Q]
    solid = ~fragile
    over = (err >= bar) & solid
    assert float(over.float().sum()) <= 0.01 * float(solid.float().sum()), '%s: %d of %d non-border rows above %.0e (worst %.3e)' % (
        what, int(over.sum()), int(solid.sum()), bar, float(err[solid].max()))
    assert float(err[solid].max()) < 10 * bar, '%s: a non-border row is %.3e off (bar %.0e, hard limit 10x)' % (what, float(err[solid].max()), bar)
    flipped = (err >= bar) & fragile
    assert float(flipped.float().mean()) <= 0.02 and float(err.max()) < 0.5, '%s: %d view-border rows differ (worst %.3e)' % (what, int(flipped.sum()), float(err.max()))
    return int(flipped.sum())


def _smooth_feats(cfg, T, seed):
    """FPN-like SMOOTH feature maps: unit-variance noise generated at 1/8 resolution and bilinearly upsampled.  The default
    synthetic maps are i.i.d. noise per pixel -- maximally rough: a sample point that moves by 1e-6 of its coordinate (the
    fp32-grade round-off of the sampling-offset Linear: tensor-core bf16x3 vs the oracle's fp32, max 2.3e-6 relative measured,
    tests/perf/parity_stages.py) changes the bilinear tap by ~2e-4 of the feature scale, which then IS the whole-layer error
    (measured 1.7e-4 at r50-T8, 3.8e-4 at vov99; the gather itself is exact to 2e-7 on identical points).  Backbone features
    are smooth at the pixel scale; on such maps the layer-level bar measures the kernels, not the roughness of the test data."""
    g = torch.Generator().manual_seed(seed)
    feats = []
    for (h, w) in cfg['levels']:
        low = torch.randn(T * 6, 256, h // 8 + 2, w // 8 + 2, generator=g)
        f = torch.nn.functional.interpolate(low, size=(h, w), mode='bilinear', align_corners=True)
        feats.append((f / f.std())[None].contiguous())
    return feats


@pytest.mark.parametrize('name,T', [('r50_704x256', 8), ('r50_704x256', 1), ('r101_1408x512', 2), ('vov99_1600x640', 2)])
def test_full_size_layer_vs_oracle(name, T):
    """The BENCH workload itself (r50 704x256, 900 queries, T = 8: BASELINE config 3's per-layer shape; T = 1: config 2)
    and the 5-level configs 4 / 5 at full resolution and query count (two frames: the CPU oracle holds the pyramid twice)
    held to the CPU oracle's restatement of SparseBEVTransformerDecoderLayer.forward
    (/root/reference/models/sparsebev_transformer.py:162-193), incl. the 900 = 7 x 128 + 4 row tail of every GEMM tile.
      * whole layer on smooth feature maps (_smooth_feats): every query row within 5e-5 of the output scale, except rows
        with a sample point on a camera's validity border (_fragile_queries);
      * every stage fed with the ORACLE's inputs (errors cannot compound): 2e-5, the gather on identical points 1e-6."""
    from sparsebev_b200 import ops
    cfg, sd, model, _, metas, qb, qf = _setup(name, T, 1, seed=1, num_layers=1)
    feats = _smooth_feats(cfg, T, seed=7)
    td = R.time_diff_from_timestamps([m['img_timestamp'] for m in metas])
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas]).astype(np.float32))
    taps = {}
    with torch.no_grad():
        want = R.decoder_layer(qb, qf, R.regroup_feats(feats, channel_last=True), sd, cfg, td, l2i, op=R.msmv_sampling_kernel_semantics, taps=taps)
        pos = qf + R.position_encoder(qb[..., :3], sd)
        ffn_want = R._ln(R.ffn(taps['mixed'], sd), sd['norm3.weight'], sd['norm3.bias'])
    fragile = _fragile_queries(taps['points'], l2i, cfg['image_h'], cfg['image_w'])
    assert float(fragile.float().mean()) < 0.25
    layer = model.decoder.decoder_layer
    metas_gpu = copy.deepcopy(metas)
    model.decoder.prepare_metas(metas_gpu, 1, torch.device('cuda'))
    gfeats = model.decoder.prepare_feats([f.cuda() for f in feats])
    got = layer(qb.cuda(), qf.cuda(), gfeats, None, metas_gpu)
    flips = 0
    for g, w, what in zip(got, want, ('query_feat', 'cls', 'bbox')):
        assert g.shape == w.shape and torch.isfinite(g).all()
        flips = max(flips, _rows_within(g, w, LAYER_BAR, fragile, '%s %s T=%d' % (name, what, T)))
    # ---- stage by stage on the oracle's inputs
    Q, D, G, P, L = cfg['num_query'], 256, 4, 4, cfg['num_levels']
    qbc = qb.cuda()
    sasa = layer.self_attn.forward_fused(qbc, pos.cuda(), None, layer.norm1)
    assert _rel(sasa, taps['after_sasa']) < 2e-5, 'SASA block rel-to-max %.3e' % _rel(sasa, taps['after_sasa'])
    after = taps['after_sasa'].cuda()
    heads = layer.sampling._heads(after.reshape(Q, D))
    pts, sw = ops.sample_points(qbc, heads, heads[:, G * P * 3:], cfg['pc_range'], L, num_points_total=G * P, ld_off=heads.shape[1], ld_log=heads.shape[1])
    want_pts = taps['points'][:, :, 0].reshape(1, Q, G * P, 3).contiguous()                  # frame 0 has time_diff 0: the un-warped points
    assert _rel(pts, want_pts) < 1e-5 and _rel(sw.reshape(1, Q, G, P, L), taps['scale_weights'][:, :, :, 0]) < 2e-5
    exact = ops.sampling4d_fused(gfeats, want_pts.cuda(), qbc, metas_gpu[0]['time_diff'], metas_gpu[0]['lidar2img'],
                                 taps['scale_weights'][:, :, :, 0].contiguous().cuda(), cfg['image_h'], cfg['image_w'], num_frames=T,
                                 layout=layer.sampling.feat_layout)
    assert _rel(exact, taps['sampled']) < 1e-6, 'gather on the oracle\'s points: rel-to-max %.3e' % _rel(exact, taps['sampled'])
    sampled = layer.sampling(qbc, after, gfeats, metas_gpu)
    _rows_within(sampled, taps['sampled'], 5e-5, fragile, 'sampled features')
    mixed = layer.mixing.forward_fused(taps['sampled'].contiguous().cuda(), after.contiguous(), layer.norm2)
    assert _rel(mixed, taps['mixed']) < 2e-5, 'mixing block rel-to-max %.3e' % _rel(mixed, taps['mixed'])
    q4 = torch.empty(Q, D, device='cuda')
    mx = taps['mixed'].cuda().reshape(Q, D).contiguous()
    ops.dense_chain(mx, D, Q, [layer._ffn0.layer(relu=True), layer._ffn1.layer(residual=mx, res_pre_ln=True, y=q4)])
    assert _rel(q4, ffn_want[0]) < 2e-5
    print('%s T=%d: %d of %d query rows are view-border cases' % (name, T, flips, Q))


def test_mix_presplit_m900_vs_oracle():
    """The production mix path (tcgen05 parameter GEMM emitting bf16 (hi, lo) -> TMA-fed mix kernel) at M = 900 rows
    directly against the oracle's fp32 restatement of AdaptiveMixing's middle stage
    (/root/reference/models/sparsebev_transformer.py:358-375), not against our own fp32-parameter kernel."""
    from sparsebev_b200 import ops
    M, G, Pin, C, D = 900, 4, 32, 64, 256
    g = torch.Generator().manual_seed(11)
    q = torch.randn(M, D, generator=g)
    W = torch.randn(G * (C * C + 128 * Pin), D, generator=g) * 0.02
    b = torch.randn(G * (C * C + 128 * Pin), generator=g) * 0.05
    x = torch.randn(M, G, Pin, C, generator=g)
    params = (q.double() @ W.double().t() + b.double()).float().reshape(M, G, -1)
    m = params[..., :C * C].reshape(M, G, C, C)
    sm = params[..., C * C:].reshape(M, G, 128, Pin)
    h = torch.relu(torch.nn.functional.layer_norm(x @ m, (Pin, C)))
    want = torch.relu(torch.nn.functional.layer_norm(sm @ h, (128, C))).reshape(M, -1)
    qh, ql = ops.split_bf16(q.cuda())
    wh, wl = ops.split_bf16(W.cuda())
    ph, pl = ops.gemm_bf16_tn_split(qh, ql, wh, wl, M, W.shape[0], D, bias=b.cuda())
    hi, lo, y = ops.mix_presplit(ph, pl, x.cuda(), want_f32=True)
    for got, what in ((y, 'fp32 output'), (hi.float() + lo.float(), 'bf16 hi+lo output')):
        err = (got.cpu() - want).abs()
        assert float(err.max() / want.abs().max()) < 2e-5, '%s: max-abs / max-ref %.3e' % (what, float(err.max() / want.abs().max()))
        # 29.5 M unit-variance outputs: the far tail of the bf16-pair operand rounding reaches 8e-5 absolute (5e-5 holds at 5 rows)
        assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-4), '%s: worst abs err %.3e' % (what, float(err.max()))
    tail = slice(896, 900)                       # the 4 valid rows of the eighth 128-row GEMM tile
    assert torch.allclose(y[tail].cpu(), want[tail], rtol=1e-4, atol=1e-4)


def test_r50_t8_layer_runs_and_is_deterministic():
    """Full-size r50-T8 (the bench workload): finite, deterministic, and the sampled block equals the
    op-boundary path (sbev_msmv_fwd fed with the fused kernel's own loc) bit for bit."""
    from sparsebev_b200 import ops
    cfg, sd, model, feats, metas, qb, qf = _setup('r50_704x256', 8, 1, seed=1, num_layers=1)
    layer = model.decoder.decoder_layer
    model.decoder.prepare_metas(metas, 1, torch.device('cuda'))
    gfeats = model.decoder.prepare_feats([f.cuda() for f in feats])
    a = layer(qb.cuda(), qf.cuda(), gfeats, None, metas)
    b = layer(qb.cuda(), qf.cuda(), gfeats, None, metas)
    for x, y in zip(a, b):
        assert torch.isfinite(x).all() and torch.equal(x, y)
    B, Q, G, P, T, Lv = 1, 900, 4, 4, 8, 4
    x = qf.cuda().reshape(Q, 256)
    heads = layer.sampling._heads(x)                                   # [Q, 48 + 64] = offset | scale logits
    pts, sw = ops.sample_points(qb.cuda(), heads, heads[:, 48:], cfg['pc_range'], Lv, num_points_total=16, ld_off=112, ld_log=112)
    pts_b, sw_b = ops.sample_points(qb.cuda(), heads[:, :48].contiguous().reshape(1, Q, 48), heads[:, 48:].contiguous().reshape(1, Q, 64),
                                    cfg['pc_range'], Lv)
    assert torch.equal(pts, pts_b) and torch.equal(sw, sw_b)           # strided and packed forms agree
    out, loc = ops.sampling4d_fused(gfeats, pts, qb[..., 8:10].contiguous().cuda(), metas[0]['time_diff'], metas[0]['lidar2img'],
                                    sw.reshape(1, Q, G, P, Lv), 256, 704, num_frames=T, return_loc=True)
    i = torch.arange(T * G, device='cuda')
    w_op = sw.reshape(1, Q, G, P, Lv)[0][:, (i // T)].permute(1, 0, 2, 3).contiguous()          # [T*G,Q,P,L], row (t*G+g) -> g'
    ref = ops.msmv_forward(gfeats, loc, w_op)                                                   # [TG,Q,C,P]
    ref = ref.reshape(1, T, G, Q, 64, P).permute(0, 3, 2, 1, 5, 4).reshape(1, Q, G, T * P, 64)
    assert torch.equal(out, ref)


@pytest.mark.parametrize('name,T', [('r101_1408x512', 8), ('vov99_1600x640', 8)])
def test_five_level_configs_full_size_properties(name, T):
    """BASELINE configs 4/5 (5 FPN levels; vov99: 1600 queries) at full size: the layer runs finite and deterministic, and the fused
    gather equals the op-boundary path (sbev_msmv_fwd fed with the fused kernel's own loc) bit for bit -- size-independent
    properties, since the CPU oracle would need minutes per layer here."""
    import sparsebev_b200 as sb
    from sparsebev_b200 import ops, synthetic as S
    cfg = S.layer_cfg(name, T, num_layers=1)
    sd = S.make_state_dict(cfg, seed=5)
    model = sb.SparseBEVTransformer(256, num_frames=T, num_points=4, num_layers=1, num_levels=cfg['num_levels'], pc_range=cfg['pc_range']).cuda().eval()
    model.load_state_dict({'decoder.decoder_layer.' + k: v for k, v in sd.items()})
    layer = model.decoder.decoder_layer
    Q, G, P, Lv = cfg['num_query'], 4, 4, cfg['num_levels']
    g = torch.Generator(device='cuda').manual_seed(7)
    feats = [torch.randn(1, T * 6, h, w, 256, device='cuda', generator=g).permute(0, 1, 4, 2, 3) for (h, w) in cfg['levels']]   # NHWC storage: zero-copy
    metas = S.make_metas(name, T, batch=1)
    model.decoder.prepare_metas(metas, 1, torch.device('cuda'))
    gfeats = model.decoder.prepare_feats(list(feats))
    assert layer.sampling.feat_layout == 'nhwc'
    qb = S.init_query_bbox(Q, seed=2)[None].cuda()
    qf = torch.randn(1, Q, 256, generator=torch.Generator().manual_seed(3)).cuda()
    a = layer(qb, qf, gfeats, None, metas)
    b = layer(qb, qf, gfeats, None, metas)
    for x, y in zip(a, b):
        assert torch.isfinite(x).all() and torch.equal(x, y)
    heads = layer.sampling._heads(qf.reshape(Q, 256))
    nh = G * P * 3
    pts, sw = ops.sample_points(qb, heads, heads[:, nh:], cfg['pc_range'], Lv, num_points_total=G * P, ld_off=heads.shape[1], ld_log=heads.shape[1])
    out, loc = ops.sampling4d_fused(gfeats, pts, qb[..., 8:10].contiguous(), metas[0]['time_diff'], metas[0]['lidar2img'],
                                    sw.reshape(1, Q, G, P, Lv), cfg['image_h'], cfg['image_w'], num_frames=T, layout='nhwc', return_loc=True)
    # op-boundary path on ONE frame (regrouping all 8 frames of vov99 would copy 4 GB): frame t's slices are rows t*G..t*G+G-1
    t = T - 1
    grouped = [f[:, t * 6:(t + 1) * 6].permute(0, 1, 3, 4, 2).reshape(1, 6, f.shape[3], f.shape[4], G, 64).permute(0, 4, 1, 2, 3, 5)
               .reshape(G, 6, f.shape[3], f.shape[4], 64).contiguous() for f in feats]
    i = torch.arange(t * G, (t + 1) * G, device='cuda')
    w_op = sw.reshape(1, Q, G, P, Lv)[0][:, (i // T)].permute(1, 0, 2, 3).contiguous()          # [G,Q,P,L]
    ref = ops.msmv_forward(grouped, loc[t * G:(t + 1) * G].contiguous(), w_op)                   # [G,Q,C,P]
    ref = ref.permute(1, 0, 3, 2)                                                                # [Q,G,P,C]
    assert torch.equal(out[0, :, :, t * P:(t + 1) * P], ref)


def test_layer_cuda_graph_mode_equals_eager():
    """`use_cuda_graph`: the layer replayed as one captured CUDA graph returns exactly what the eager launches return
    (same kernels, same order) -- for changing query inputs (device or pinned-host sources), for the whole 3-layer decoder
    built on it, and it re-captures when the feature maps move or a kernel-variant option changes."""
    from sparsebev_b200 import _lib
    B, L = 2, 3
    cfg, sd, model, feats, metas, qb, qf = _setup('tiny', 4, B, seed=21, num_layers=L)
    layer = model.decoder.decoder_layer
    metas_gpu = copy.deepcopy(metas)
    model.decoder.prepare_metas(metas_gpu, B, torch.device('cuda'))
    gfeats = model.decoder.prepare_feats([f.cuda() for f in feats])
    inputs = [(qb, qf), (qb.flip(1).contiguous(), 0.5 * qf), (qb, qf)]
    eager = [[t.clone() for t in layer(a.cuda(), b.cuda(), gfeats, None, metas_gpu)] for a, b in inputs]
    layer.use_cuda_graph = True
    try:
        for i, (a, b) in enumerate(inputs):
            src = (a.pin_memory(), b.pin_memory()) if i == 1 else (a.cuda(), b.cuda())     # host (pinned) sources are accepted too
            got = layer(src[0], src[1], gfeats, None, metas_gpu)
            torch.cuda.synchronize()
            for g, e in zip(got, eager[i]):
                assert torch.equal(g, e), 'graph replay %d differs from the eager launches' % i
        assert len(layer._graphs) == 1
        moved = [f.clone() for f in gfeats]                                  # same values at new addresses -> a second graph
        got = layer(qb.cuda(), qf.cuda(), moved, None, metas_gpu)
        assert len(layer._graphs) == 2 and all(torch.equal(g, e) for g, e in zip(got, eager[0]))
        _lib.set_option('pdl', _lib.get_option('pdl'))                       # any set_option bumps the epoch -> re-capture
        layer(qb.cuda(), qf.cuda(), gfeats, None, metas_gpu)
        assert len(layer._graphs) == 3
        # whole decoder (shared-weight layer looped L times, reference :86-99) through the graphed layer
        layer.use_cuda_graph = False
        want = model(qb.cuda(), qf.cuda(), [f.cuda() for f in feats], None, copy.deepcopy(metas))
        layer.use_cuda_graph = True
        got = model(qb.cuda(), qf.cuda(), [f.cuda() for f in feats], None, copy.deepcopy(metas))
        assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    finally:
        layer.use_cuda_graph = False
        layer.reset_graphs()


def test_decoder_cuda_graph_mode_equals_eager():
    """`decoder.use_cuda_graph`: ALL layers of a forward captured into one CUDA graph (inputs copied in and the two stacked results
    copied out once per forward) return exactly what the eager launches return, for changing queries, and the graph is captured once
    while the (channels-last, consumed in place) feature maps keep their addresses."""
    B, L = 1, 3
    cfg, sd, model, feats, metas, qb, qf = _setup('tiny', 4, B, seed=23, num_layers=L, memory_format='nhwc')
    layer = model.decoder.decoder_layer
    gf = [f.cuda() for f in feats]
    inputs = [(qb, qf), (qb.flip(1).contiguous(), 0.5 * qf)]
    want = [[t.clone() for t in model(a.cuda(), b.cuda(), list(gf), None, copy.deepcopy(metas))] for a, b in inputs]
    assert want[0][0].shape[0] == L and not torch.equal(want[0][0], want[1][0])
    model.decoder.use_cuda_graph = True
    try:
        for i, (a, b) in enumerate(inputs * 2):
            got = model(a.cuda(), b.cuda(), list(gf), None, copy.deepcopy(metas))
            torch.cuda.synchronize()
            assert torch.equal(got[0], want[i % 2][0]) and torch.equal(got[1], want[i % 2][1]), 'decoder graph replay %d differs from eager' % i
        assert len(layer._graphs) == 1
    finally:
        model.decoder.use_cuda_graph = False
        layer.reset_graphs()

