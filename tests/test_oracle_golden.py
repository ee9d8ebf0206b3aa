"""Pin the CPU oracle (oracle/ref_torch.py) against golden vectors produced by the REAL
reference (oracle/gen_golden.py, run in the build container where /root/reference exists)."""
import os

import numpy as np
import torch

from oracle import ref_torch as R
from oracle.synth import hashrand


def _load(golden_dir, name):
    return {k: v for k, v in np.load(os.path.join(golden_dir, name)).items()}


def _feats(g, channel_first=True):
    Bp, C, N = [int(v) for v in g['shape'][:3]] if 'shape' in g else (None, None, None)
    return Bp, C, N


def test_geometry_matches_reference(golden_dir):
    g = _load(golden_dir, 'geometry.npz')
    boxes = torch.from_numpy(g['boxes'])
    dec = R.decode_bbox(boxes, g['pc_range'].tolist())
    np.testing.assert_allclose(dec.numpy(), g['decoded'], rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(R.inverse_sigmoid(boxes[..., 0:3]).numpy(), g['inv_sig'], rtol=1e-6, atol=1e-6)


def _op_case(golden_dir, name):
    g = _load(golden_dir, name)
    Bp, C, N, Q, P = [int(v) for v in g['shape']]
    feats_cf = [hashrand((Bp, C, N, int(h), int(w)), int(s), -1.0, 1.0)
                for (h, w), s in zip(g['hw'], g['feat_seeds'])]
    return g, feats_cf, torch.from_numpy(g['loc']), torch.from_numpy(g['w'])


def test_op_gridsample_restatement_is_reference(golden_dir):
    for name in ('op_small.npz', 'op_cfg1.npz'):
        g, feats, loc, w = _op_case(golden_dir, name)
        out = R.msmv_sampling_gridsample(feats, loc, w)
        np.testing.assert_allclose(out.numpy(), g['out'], rtol=0, atol=1e-6)


def test_op_kernel_semantics_close_to_reference_gridsample(golden_dir):
    """The CUDA-kernel semantics (rounded view, 2-D bilinear) and the grid_sample path agree
    whenever the view coordinate is an exact integer/(N-1): SURVEY.md section 0, gotcha 3."""
    for name in ('op_small.npz', 'op_cfg1.npz'):
        g, feats, loc, w = _op_case(golden_dir, name)
        feats_cl = [f.permute(0, 2, 3, 4, 1).contiguous() for f in feats]
        out = R.msmv_sampling_kernel_semantics(feats_cl, loc, w)
        np.testing.assert_allclose(out.numpy(), g['out'], rtol=1e-4, atol=2e-5)


def test_sampling_4d_matches_reference(golden_dir):
    g = _load(golden_dir, 'sampling4d.npz')
    B, Q, T, G, P, L, C, ih, iw = [int(v) for v in g['dims']]
    pc = g['pc_range'].tolist()
    qb = torch.from_numpy(g['query_bbox'])
    pts = R.make_sample_points(qb, torch.from_numpy(g['offset']), pc)
    np.testing.assert_allclose(pts.numpy(), g['points'], rtol=1e-5, atol=1e-5)
    pts6 = torch.from_numpy(g['points6'])
    sw = torch.from_numpy(g['scale_weights'])
    l2i = torch.from_numpy(g['lidar2img'])
    loc, w = R.sampling_4d(pts6, None, sw, l2i, ih, iw, return_loc=True)
    np.testing.assert_array_equal(w.numpy(), g['w'])
    # view choice must be identical; uv to fp32 round-off of the 4x4 mat-vec
    np.testing.assert_array_equal(loc[..., 2].numpy(), g['loc'][..., 2])
    np.testing.assert_allclose(loc[..., :2].numpy(), g['loc'][..., :2], rtol=1e-5, atol=1e-5)
    feats = [hashrand((B * T * G, C, 6, int(h), int(w_)), int(s), -1.0, 1.0)
             for (h, w_), s in zip(g['hw'], g['feat_seeds'])]
    out = R.sampling_4d(pts6, feats, sw, l2i, ih, iw)
    np.testing.assert_allclose(out.numpy(), g['out'], rtol=1e-4, atol=1e-4)
    # the (t,g) vs (g,t) flattening quirk is really there (gotcha 2)
    assert not np.array_equal(g['w'].reshape(B, T, G, Q, P, L), g['w'].reshape(B, G, T, Q, P, L).transpose(0, 2, 1, 3, 4, 5)) or T == 1


def test_adaptive_mixing_matches_reference(golden_dir):
    g = _load(golden_dir, 'mixing.npz')
    Bm, Qm, G, Pin, C = [int(v) for v in g['dims']]
    s = [int(v) for v in g['seeds']]
    sd = {
        'mixing.parameter_generator.weight': hashrand((G * (C * C + Pin * 128), 256), s[0], -0.04, 0.04),
        'mixing.parameter_generator.bias': hashrand((G * (C * C + Pin * 128),), s[1], -0.1, 0.1),
        'mixing.out_proj.weight': hashrand((256, G * 128 * C), s[2], -0.02, 0.02),
        'mixing.out_proj.bias': hashrand((256,), s[3], -0.05, 0.05),
    }
    x = hashrand((Bm, Qm, G, Pin, C), s[4], -2, 2)
    q = hashrand((Bm, Qm, 256), s[5], -1.5, 1.5)
    out = R.adaptive_mixing(x, q, sd)
    np.testing.assert_allclose(out.numpy(), g['out'], rtol=1e-5, atol=1e-5)


def test_self_attention_matches_torch_mha():
    """mmcv 1.6.0 MultiheadAttention wraps nn.MultiheadAttention and adds the identity; pin our
    restatement against torch's own module (the part that is available here)."""
    torch.manual_seed(0)
    B, Q, D = 2, 17, 256
    mha = torch.nn.MultiheadAttention(D, 8, dropout=0.1).eval()
    qf = torch.randn(B, Q, D)
    qb = R.init_query_bbox(25, 1)[:Q][None].repeat(B, 1, 1)
    pc = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    sd = {'self_attn.attention.attn.in_proj_weight': mha.in_proj_weight.detach(),
          'self_attn.attention.attn.in_proj_bias': mha.in_proj_bias.detach(),
          'self_attn.attention.attn.out_proj.weight': mha.out_proj.weight.detach(),
          'self_attn.attention.attn.out_proj.bias': mha.out_proj.bias.detach(),
          'self_attn.gen_tau.weight': torch.randn(8, D) * 0.05,
          'self_attn.gen_tau.bias': torch.rand(8) * 2}
    ours = R.scale_adaptive_self_attention(qb, qf, sd, pc)
    tau = torch.nn.functional.linear(qf, sd['self_attn.gen_tau.weight'], sd['self_attn.gen_tau.bias'])
    dist = -R.pairwise_centre_dist(qb, pc)
    mask = (dist[:, None] * tau.permute(0, 2, 1)[..., None]).flatten(0, 1)
    with torch.no_grad():
        ref = qf + mha(qf.transpose(0, 1), qf.transpose(0, 1), qf.transpose(0, 1), attn_mask=mask)[0].transpose(0, 1)
    np.testing.assert_allclose(ours.numpy(), ref.numpy(), rtol=1e-4, atol=1e-5)


def test_nms_free_coder_matches_reference_golden(golden_dir):
    """Post-processing mirror vs the REAL reference NMSFreeCoder (oracle/gen_golden_coder.py): exact."""
    from sparsebev_b200.coder import NMSFreeCoder
    g = np.load(os.path.join(golden_dir, 'coder.npz'))
    cls, box = torch.from_numpy(g['cls']), torch.from_numpy(g['box'])
    post = [float(v) for v in g['post_center_range']]
    for tag, kw in [('a', dict(max_num=100, score_threshold=None)), ('b', dict(max_num=37, score_threshold=0.05))]:
        coder = NMSFreeCoder(pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], post_center_range=post, num_classes=10, **kw)
        res = coder.decode({'all_cls_scores': cls, 'all_bbox_preds': box})
        assert len(res) == cls.shape[1]
        for b, r in enumerate(res):
            assert np.array_equal(r['labels'].numpy(), g['%s%d_labels' % (tag, b)])
            assert np.array_equal(r['scores'].numpy(), g['%s%d_scores' % (tag, b)])
            assert np.array_equal(r['bboxes'].numpy(), g['%s%d_bboxes' % (tag, b)])


def test_op_backward_matches_reference_autograd(golden_dir):
    """Backward of the op: the C restatement of the reference CUDA backward kernel's arithmetic
    (oracle/msmv_oracle.c) against the gradients autograd derives through the REAL reference's msmv_sampling_pytorch
    (oracle/gen_golden_bwd.py).  This pins the checker the CUDA backward kernels -- atomic and deterministic -- are compared
    with on the GPU.  The view-coordinate gradient is excluded: the reference CUDA kernel leaves it at zero."""
    from oracle import c_oracle
    g = _load(golden_dir, 'op_bwd.npz')
    Bp, C, N, Q, P = [int(v) for v in g['shape']]
    feats_cl = [hashrand((Bp, C, N, int(h), int(w)), int(s), -1.0, 1.0).permute(0, 2, 3, 4, 1).contiguous()
                for (h, w), s in zip(g['hw'], g['feat_seeds'])]
    loc, w, go = torch.from_numpy(g['loc']), torch.from_numpy(g['w']), torch.from_numpy(g['grad_out'])
    np.testing.assert_allclose(c_oracle.fwd(feats_cl, loc, w).numpy(), g['out'], rtol=1e-4, atol=2e-5)
    gf, gl, gw = c_oracle.bwd(go, feats_cl, loc, w)
    for i, a in enumerate(gf):
        np.testing.assert_allclose(a.numpy(), g['grad_feat%d' % i], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gw.numpy(), g['grad_w'], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(gl[..., :2].numpy(), g['grad_loc_uv'], rtol=1e-4, atol=2e-4)
    assert float(gl[..., 2].abs().max()) == 0.0


def test_decoder_restatement_matches_real_reference_decoder(golden_dir):
    """oracle/ref_torch.py::decoder against the REAL reference decoder (sparsebev_transformer.py executed unmodified on the
    CPU by oracle/gen_golden_decoder.py with only mmcv's MultiheadAttention / FFN / BaseModule and mmdet's registry stubbed):
    multi-layer outputs incl. box refinement, velocity rescale, metadata handling and the feature regroup.  Two cases:
    2 levels / T=2 / B=2 / 3 layers, 5 levels / T=3 / B=1 / 2 layers, and a query-denoising style attention mask (2 layers).  Also run with the CUDA-kernel sampling semantics
    (rounded view, 2-D bilinear), which is what the GPU tests compare the CUDA decoder with."""
    import pytest
    from sparsebev_b200 import synthetic as S
    g = _load(golden_dir, 'decoder.npz')
    for tag, name in zip(('a', 'b', 'c'), g['names']):
        name = str(name)
        T, B, L = [int(v) for v in g[tag + '_cfg']]
        cfg = S.layer_cfg(name, T, num_layers=L)
        sd = S.make_state_dict(cfg, seed=11)
        feats = S.make_feats(name, T, batch=B, seed=12)
        check = np.array([float(feats[0].double().sum()), float(sd['mixing.out_proj.weight'].double().sum())])
        if not np.allclose(check, g[tag + '_check'], rtol=0, atol=1e-6):
            pytest.skip('torch CPU generator stream differs from the one the fixture was generated with')
        metas = S.make_metas(name, T, batch=B)
        td = R.time_diff_from_timestamps([m['img_timestamp'] for m in metas])
        l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas]).astype(np.float32))
        qb, qf = torch.from_numpy(g[tag + '_qb']), torch.from_numpy(g[tag + '_qf'])
        mask = torch.from_numpy(g[tag + '_mask']) if (tag + '_mask') in g else None      # case c: query-denoising attention mask
        with torch.no_grad():
            cls, box = R.decoder(qb, qf, feats, sd, cfg, td, l2i, pre_attn_mask=mask)   # grid_sample path, like the reference on CPU
            cls_k, box_k = R.decoder(qb, qf, feats, sd, cfg, td, l2i, pre_attn_mask=mask, channel_last=True, op=R.msmv_sampling_kernel_semantics)
        np.testing.assert_allclose(cls.numpy(), g[tag + '_cls'], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(box.numpy(), g[tag + '_box'], rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(cls_k.numpy(), g[tag + '_cls'], rtol=1e-3, atol=2e-4)
        np.testing.assert_allclose(box_k.numpy(), g[tag + '_box'], rtol=1e-3, atol=2e-4)


def test_head_restatement_and_mirror_match_real_reference_head(golden_dir):
    """SparseBEVHead eval path against the REAL reference head run on the CPU by oracle/gen_golden_head.py
    (sparsebev_head.py unmodified; mmdet's DETRHead replaced by an arithmetic-free stand-in):
      * oracle/ref_torch.py::head_forward (query init -> decoder -> metres + reorder) reproduces all_cls_scores / all_bbox_preds;
      * the product mirror's `_init_layers` lays the learned query boxes out like the reference (grid xy, z = 0, h = 1.5, v = 0);
      * the product mirror's `get_bboxes` (+ NMSFreeCoder mirror) turns the reference's predictions into the reference's detections."""
    import pytest
    import sparsebev_b200 as sb
    from sparsebev_b200 import synthetic as S
    g = _load(golden_dir, 'head.npz')
    T, B, L, Q = [int(v) for v in g['cfg']]
    cfg = S.layer_cfg('tiny', T, num_layers=L)
    sd = S.make_state_dict(cfg, seed=21)
    feats = S.make_feats('tiny', T, batch=B, seed=22)
    check = np.array([float(feats[0].double().sum()), float(sd['mixing.out_proj.weight'].double().sum())])
    if not np.allclose(check, g['check'], rtol=0, atol=1e-6):
        pytest.skip('torch CPU generator stream differs from the one the fixture was generated with')
    metas = S.make_metas('tiny', T, batch=B)
    td = R.time_diff_from_timestamps([m['img_timestamp'] for m in metas])
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas]).astype(np.float32))
    with torch.no_grad():
        out = R.head_forward(torch.from_numpy(g['init_query_bbox']), torch.from_numpy(g['label_enc']), feats, sd, cfg, td, l2i)
    np.testing.assert_allclose(out['all_cls_scores'].numpy(), g['all_cls_scores'], rtol=1e-4, atol=2e-5)
    np.testing.assert_allclose(out['all_bbox_preds'].numpy(), g['all_bbox_preds'], rtol=1e-4, atol=1e-4)

    post = g['post_center_range'].tolist()
    head = sb.SparseBEVHead(num_classes=10, in_channels=256, num_query=Q, pc_range=cfg['pc_range'],
                            bbox_coder=dict(type='NMSFreeCoder', pc_range=cfg['pc_range'], post_center_range=post, max_num=20,
                                            score_threshold=None, num_classes=10),
                            transformer=dict(type='SparseBEVTransformer', embed_dims=256, num_frames=T, num_points=4, num_layers=L,
                                             num_levels=cfg['num_levels'], pc_range=cfg['pc_range']))
    w = head.init_query_bbox.weight.detach().numpy()
    for cols in ([0, 1], [2], [5], [8, 9]):
        np.testing.assert_array_equal(w[:, cols], g['init_query_bbox'][:, cols])
    assert tuple(head.label_enc.weight.shape) == g['label_enc'].shape
    dets = head.get_bboxes({'all_cls_scores': torch.from_numpy(g['all_cls_scores']).clone(),
                            'all_bbox_preds': torch.from_numpy(g['all_bbox_preds']).clone()})
    assert len(dets) == B
    for b, (boxes, scores, labels) in enumerate(dets):
        np.testing.assert_allclose(boxes.numpy(), g['det%d_boxes' % b], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(scores.numpy(), g['det%d_scores' % b], rtol=1e-6, atol=1e-7)
        np.testing.assert_array_equal(labels.numpy(), g['det%d_labels' % b])
