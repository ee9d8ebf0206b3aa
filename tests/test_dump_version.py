"""DUMP export hook and legacy VERSION switch (reference: models/utils.py:309-325) -- host-side / torch parts on the CPU."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_torch as R


def _inputs(B=2, Q=7, T=3, GP=8, seed=0):
    from sparsebev_b200 import synthetic as S
    g = torch.Generator().manual_seed(seed)
    pts = torch.randn(B, Q, GP, 3, generator=g) * 20
    vel = torch.randn(B, Q, 2, generator=g)
    td = (torch.arange(T, dtype=torch.float32) * 0.5)[None].repeat(B, 1)
    l2i, _ = S.camera_rig(T, 256, 704)
    return pts, vel, td, l2i[None].repeat(B, 1, 1, 1).contiguous()


def test_projected_sample_points_matches_oracle_projection():
    """The slow export path reproduces what sampling_4d projects (all T x N views, eps-clamped depth, validity)."""
    from sparsebev_b200.transformer import projected_sample_points
    pts, vel, td, l2i = _inputs()
    B, Q, GP, _ = pts.shape
    T = td.shape[1]
    cam, valid = projected_sample_points(pts, vel, td, l2i, 256, 704)
    warped = pts[:, :, None].expand(B, Q, T, GP, 3).clone()
    warped[..., :2] -= (vel[:, :, None, :] * td[:, None, :, None])[:, :, :, None, :]
    _, _, want_cam, want_valid = R.project_and_select_view(warped, l2i, 256, 704, return_all=True)
    assert cam.shape == (B, T, 6, Q, GP, 3) and valid.shape == (B, T, 6, Q, GP)
    assert torch.allclose(cam, want_cam, rtol=1e-4, atol=1e-4)          # (points behind a camera divide by eps: huge values, relative bar)
    assert (valid.bool() == want_valid).float().mean() > 0.999          # einsum vs fixed-order mat-vec: a border point may flip
    assert 0.02 < float(valid.mean()) < 0.5


@pytest.mark.skipif(not os.path.isdir('/root/reference/models'), reason='the real reference is only present in the build container')
def test_dump_files_match_the_real_reference(tmp_path):
    """Same file contents as the REAL reference's sampling_4d writes under DUMP.enabled (models/sparsebev_sampling.py:82-86)."""
    from oracle.gen_golden import import_reference
    from sparsebev_b200.transformer import projected_sample_points
    _, utils, _, sampling, _ = import_reference()
    pts, vel, td, l2i = _inputs(B=1, Q=5, T=2, GP=16, seed=3)
    B, Q, GP, _ = pts.shape
    T, G, P = td.shape[1], 4, 4
    warped = pts[:, :, None].expand(B, Q, T, GP, 3).clone()
    warped[..., :2] -= (vel[:, :, None, :] * td[:, None, :, None])[:, :, :, None, :]
    feats = [torch.randn(B * T * G, 8, 6, 4, 5)]
    sw = torch.ones(B, Q, G, T, P, 1)
    utils.DUMP.enabled, utils.DUMP.out_dir, utils.DUMP.stage_count = True, str(tmp_path), 0
    try:
        sampling.sampling_4d(warped.reshape(B, Q, T, G, P, 3), feats, sw, l2i, 256, 704)
    finally:
        utils.DUMP.enabled = False
    want_cam = torch.load(os.path.join(str(tmp_path), 'sample_points_cam_stage0.pth'))
    want_valid = torch.load(os.path.join(str(tmp_path), 'sample_points_cam_valid_mask_stage0.pth'))
    cam, valid = projected_sample_points(pts, vel, td, l2i, 256, 704)
    assert cam.shape == want_cam.shape and valid.shape == want_valid.shape and valid.dtype == want_valid.dtype
    assert torch.allclose(cam, want_cam, rtol=1e-4, atol=1e-4)
    agree = (valid == want_valid).float().mean()
    assert agree > 0.999          # (a point within 1 ulp of the image border may flip)


def test_decode_bbox_mirror_and_version_switch(monkeypatch):
    from sparsebev_b200 import _lib, utils
    from sparsebev_b200.transformer import decode_bbox
    b = torch.randn(3, 5, 10)
    assert torch.allclose(decode_bbox(b, [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]), R.decode_bbox(b, [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]), atol=1e-6)
    calls = []
    state = {'legacy_rotation': 0}
    monkeypatch.setattr(_lib, 'get_option', lambda name: state[name])
    monkeypatch.setattr(_lib, 'set_option', lambda name, v: (calls.append((name, v)), state.__setitem__(name, v)))
    v = utils.Version()
    assert v.name == 'v1.0.0' and not v.legacy
    v.name = 'v0.17.1'
    assert v.legacy and calls == [('legacy_rotation', 1)]
    v.name = 'v0.17.1'
    assert len(calls) == 1                                   # no redundant option writes (they invalidate captured graphs)
    v.name = 'v1.0.0'
    assert calls[-1] == ('legacy_rotation', 0)
    monkeypatch.setattr(utils, 'VERSION', v)
    assert utils.load_checkpoint_version({'version': 'v0.17.1', 'state_dict': {}}) == 'v0.17.1' and v.legacy
    assert utils.load_checkpoint_version({'state_dict': {}}) == 'v0.17.1'          # absent key: unchanged, like val.py:128-129


def test_head_get_bboxes_legacy_swap(monkeypatch):
    """models/sparsebev_head.py:472-476: with VERSION 'v0.17.1' w / l are swapped and yaw -> -yaw - pi/2."""
    import math
    import sparsebev_b200 as sb
    from sparsebev_b200 import head as H, utils
    pc = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    hd = sb.SparseBEVHead(num_classes=10, in_channels=256, num_query=4, pc_range=pc,
                          transformer=dict(type='SparseBEVTransformer', embed_dims=256, num_frames=1, num_points=4, num_layers=1, num_levels=1, pc_range=pc),
                          bbox_coder=dict(type='NMSFreeCoder', pc_range=pc, post_center_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], max_num=4, num_classes=10))
    g = torch.Generator().manual_seed(1)
    preds = {'all_cls_scores': torch.randn(1, 1, 4, 10, generator=g), 'all_bbox_preds': torch.randn(1, 1, 4, 10, generator=g)}
    base = hd.get_bboxes({k: v.clone() for k, v in preds.items()})[0]

    class V:
        legacy = True
    monkeypatch.setattr(utils, 'VERSION', V())
    leg = hd.get_bboxes({k: v.clone() for k, v in preds.items()})[0]
    assert torch.equal(leg[1], base[1]) and torch.equal(leg[2], base[2])
    assert torch.equal(leg[0][:, 3], base[0][:, 4]) and torch.equal(leg[0][:, 4], base[0][:, 3])
    assert torch.allclose(leg[0][:, 6], -base[0][:, 6] - math.pi / 2)
    assert torch.equal(leg[0][:, [0, 1, 2, 5, 7, 8]], base[0][:, [0, 1, 2, 5, 7, 8]])


def test_invalidate_caches_hooks():
    """Derived weight copies / captured graphs are dropped after init_weights, load_state_dict and .to() -- the paths that can
    change parameter values without bumping the (data_ptr, _version) cache key (`.data` writes)."""
    import sparsebev_b200 as sb
    m = sb.SparseBEVTransformer(256, num_frames=2, num_points=4, num_layers=1, num_levels=2, pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0])
    layer = m.decoder.decoder_layer
    caches = [layer._pe0.cache, layer.self_attn._cache_in, layer.mixing._pg, layer.sampling._heads.cache]

    def poison():
        for c in caches:
            c.key = 'stale'
        layer._graphs['x'] = None
    poison(); m.init_weights()
    assert all(c.key is None for c in caches) and len(layer._graphs) == 0
    poison(); m.load_state_dict(m.state_dict())
    assert all(c.key is None for c in caches) and len(layer._graphs) == 0
    poison(); m.double()
    assert all(c.key is None for c in caches) and len(layer._graphs) == 0
    poison(); layer.invalidate_caches()
    assert all(c.key is None for c in caches)


def test_decoder_level_graph_dispatch(monkeypatch):
    """decoder.use_cuda_graph sends a forward through ONE graph (layer._forward_graphed with the all-layers body) -- except with the DUMP
    export on or a sharded layer, which take the plain loop -- and the all-layers body runs with the per-layer graphs switched off."""
    import sparsebev_b200 as sb
    from sparsebev_b200.utils import DUMP
    m = sb.SparseBEVTransformer(256, num_frames=2, num_points=4, num_layers=3, num_levels=2, pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0])
    dec, layer = m.decoder, m.decoder.decoder_layer
    monkeypatch.setattr(dec, 'prepare_metas', lambda *a, **k: None)
    monkeypatch.setattr(dec, 'prepare_feats', lambda f: f)
    monkeypatch.setattr(torch.cuda, 'is_current_stream_capturing', lambda: False)
    calls = []

    def fake_layer(qb, qf, feats, mask, metas):
        calls.append(('layer', layer.use_cuda_graph))
        return qf, torch.zeros(1, 4, 10), torch.zeros(1, 4, 10)
    monkeypatch.setattr(layer, 'forward', fake_layer)

    def fake_graphed(qb, qf, feats, mask, metas, impl=None):
        calls.append(('graphed', impl.__name__))
        return impl(qb, qf, feats, mask, metas)
    monkeypatch.setattr(layer, '_forward_graphed', fake_graphed)
    args = (torch.zeros(1, 4, 10), torch.zeros(1, 4, 256), [], None, [{}])
    cls, box = dec(*args)
    assert cls.shape == (3, 1, 4, 10) and [c[0] for c in calls] == ['layer'] * 3
    calls.clear()
    dec.use_cuda_graph, layer.use_cuda_graph = True, True
    dec(*args)
    assert calls[0] == ('graphed', '_layers_one_graph') and calls[1:] == [('layer', False)] * 3      # per-layer graphs off inside the capture
    assert layer.use_cuda_graph is True                                                             # ... and restored afterwards
    calls.clear()
    monkeypatch.setattr(DUMP, 'enabled', True)
    dec(*args)
    assert [c[0] for c in calls] == ['layer'] * 3
    monkeypatch.setattr(DUMP, 'enabled', False)
    calls.clear()

    class _Shard:
        world = 2
    layer.query_shard = _Shard()
    dec(*args)
    assert [c[0] for c in calls] == ['layer'] * 3
