"""GPU parity tests: every CUDA kernel (called through the C ABI) against the CPU oracle on the same
seeded inputs, against the committed golden vectors of the real reference, and -- when the compiled
reference op travelled with the snapshot (oracle/_ref) -- against the reference CUDA kernel itself.

Tolerances (stated per north_star): sample INDICES bit-exact; sampled / mixed FEATURES within
1e-4 relative with a 1e-5 absolute floor (gotcha 3).
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle, ref_torch as R
from oracle.synth import hashrand

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-5


def dev():
    return torch.device('cuda:0')


def _ops():
    from sparsebev_b200 import ops
    return ops


@pytest.fixture
def option():
    """Select a kernel variant for one test, restore the defaults afterwards."""
    from sparsebev_b200 import _lib
    names = ('gemm_impl', 'mix_impl', 'sasa_impl', 'gather_variant', 'dense_impl', 'dense_cluster', 'dense_nsplit', 'dense_pack', 'legacy_rotation', 'sasa_kq',
             'dense_ws', 'dense_ws_groups')
    defaults = {k: _lib.get_option(k) for k in names}

    def setter(name, value):
        _lib.set_option(name, value)
    yield setter
    for k, v in defaults.items():
        _lib.set_option(k, v)


def _close(a, b, rtol=RTOL, atol=ATOL, what=''):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), '%s: %d / %d elements off, max abs err %.3e (max ref %.3e)' % (
        what, int(bad.sum()), bad.numel(), float(err.max()), float(b.abs().max()))


def _rand_case(Bp, N, hw, Q, P, C=64, seed=0, spread=0.15):
    feats = [hashrand((Bp, N, h, w, C), seed * 10 + i, -1, 1) for i, (h, w) in enumerate(hw)]
    loc = hashrand((Bp, Q, P, 3), seed * 10 + 7, -spread, 1 + spread)
    loc[..., 2] = torch.from_numpy(np.random.RandomState(seed).randint(0, N, size=(Bp, Q, P))).float() / max(N - 1, 1)
    w = torch.softmax(hashrand((Bp, Q, P, len(hw)), seed * 10 + 8, -2, 2), dim=-1)
    return feats, loc, w


def _ref_cuda():
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref', '_msmv_sampling_cuda.so')
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location('_msmv_sampling_cuda', so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


# ------------------------------------------------------------------------------------------ op
@pytest.mark.parametrize('name', ['op_small.npz', 'op_cfg1.npz'])
def test_msmv_fwd_matches_golden_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name))
    Bp, C, N, Q, P = [int(v) for v in g['shape']]
    feats = [hashrand((Bp, C, N, int(h), int(w)), int(s), -1, 1).permute(0, 2, 3, 4, 1).contiguous().to(dev())
             for (h, w), s in zip(g['hw'], g['feat_seeds'])]
    out = _ops().msmv_forward(feats, torch.from_numpy(g['loc']).to(dev()), torch.from_numpy(g['w']).to(dev()))
    _close(out, torch.from_numpy(g['out']), what='vs reference msmv_sampling_pytorch (%s)' % name)


@pytest.mark.parametrize('L,P,Q,C', [(4, 4, 37, 64), (5, 4, 20, 64), (1, 1, 9, 64), (2, 7, 11, 64), (3, 32, 5, 64),
                                     (4, 4, 13, 16), (2, 3, 8, 8)])
def test_msmv_fwd_vs_c_oracle(L, P, Q, C):
    hw = [(12, 20), (6, 10), (3, 5), (2, 3), (1, 2)][:L]
    feats, loc, w = _rand_case(3, 6, hw, Q, P, C=C, seed=L * 100 + P)
    want = c_oracle.fwd(feats, loc, w)
    got = _ops().msmv_forward([f.to(dev()) for f in feats], loc.to(dev()), w.to(dev()))
    _close(got, want, what='fwd L=%d P=%d C=%d' % (L, P, C))


def test_msmv_indices_bit_exact():
    hw = [(64, 176), (32, 88), (16, 44), (8, 22)]
    loc = hashrand((4, 900, 4, 3), 5, -0.2, 1.2)
    loc[..., 2] = torch.from_numpy(np.random.RandomState(1).randint(0, 6, size=(4, 900, 4))).float() / 5
    loc[0, :8, 0, 0] = torch.tensor([0.0, 1.0, 0.5, 1 / 175, 174.5 / 175, -1 / 175, 1 + 1 / 175, 0.999999])
    want = c_oracle.indices(hw, loc, 6)
    got = _ops().msmv_indices(hw, loc.to(dev()), 6)
    for name, a, b in zip(('view', 'y0', 'x0', 'inside'), got, want):
        assert torch.equal(a.cpu(), b), name + ' differs from the reference index arithmetic'


def test_msmv_fwd_r50_t1_full_size_properties():
    """BASELINE config 2 shape (r50 704x256, 4 levels, 6 cams, T=1 -> B'=4, 900 q): checked through
    size-independent properties: linearity in the features and in the scale weights, zero outside."""
    hw = [(64, 176), (32, 88), (16, 44), (8, 22)]
    torch.manual_seed(0)
    feats = [torch.randn(4, 6, h, w, 64, device=dev()) for h, w in hw]
    loc = torch.rand(4, 900, 4, 3, device=dev())
    loc[..., 2] = torch.randint(0, 6, (4, 900, 4), device=dev()).float() / 5
    w = torch.softmax(torch.randn(4, 900, 4, 4, device=dev()), -1)
    ops = _ops()
    a = ops.msmv_forward(feats, loc, w)
    b = ops.msmv_forward([2.5 * f for f in feats], loc, w)
    _close(b, 2.5 * a, what='linearity in features')
    c = ops.msmv_forward(feats, loc, 0.5 * w)
    _close(c, 0.5 * a, what='linearity in weights')
    far = loc.clone()
    far[..., 0] = 3.0
    assert float(ops.msmv_forward(feats, far, w).abs().max()) == 0.0
    # sub-sample against the oracle (first batch slice, first 40 queries)
    want = c_oracle.fwd([f[:1].cpu() for f in feats], loc[:1, :40].cpu(), w[:1, :40].cpu())
    _close(a[:1, :40], want, what='r50-T1 subsample vs oracle')


def test_msmv_empty_and_errors():
    ops = _ops()
    feats = [torch.zeros(2, 6, 4, 4, 64, device=dev())]
    out = ops.msmv_forward(feats, torch.zeros(2, 0, 4, 3, device=dev()), torch.zeros(2, 0, 4, 1, device=dev()))
    assert out.shape == (2, 0, 64, 4)
    with pytest.raises(RuntimeError, match='num_point exceed limits'):      # reference: msmv_sampling.cpp:125
        ops.msmv_forward(feats, torch.zeros(2, 1, 33, 3, device=dev()), torch.zeros(2, 1, 33, 1, device=dev()))
    with pytest.raises(RuntimeError, match='contiguous'):                    # reference: msmv_sampling.cpp:106-111
        ops.msmv_forward([feats[0].transpose(2, 3)], torch.zeros(2, 1, 4, 3, device=dev()), torch.zeros(2, 1, 4, 1, device=dev()))
    with pytest.raises(RuntimeError, match='CUDA'):
        ops.msmv_forward(feats, torch.zeros(2, 1, 4, 3), torch.zeros(2, 1, 4, 1, device=dev()))


@pytest.mark.parametrize('L,P,C', [(4, 4, 64), (2, 3, 64), (5, 4, 64), (2, 2, 8)])
def test_msmv_bwd_vs_c_oracle(L, P, C):
    hw = [(7, 9), (4, 6), (3, 3), (2, 2), (1, 2)][:L]
    feats, loc, w = _rand_case(2, 6, hw, 10, P, C=C, seed=40 + L)
    go = hashrand((2, 10, C, P), 99, -1, 1)
    gf, gl, gw = c_oracle.bwd(go, feats, loc, w)
    f2, l2, w2 = _ops().msmv_backward(go.to(dev()), [f.to(dev()) for f in feats], loc.to(dev()), w.to(dev()))
    for a, b in zip(f2, gf):
        _close(a, b, atol=1e-4, what='grad feats')
    _close(l2, gl, atol=1e-3, what='grad loc')
    _close(w2, gw, atol=1e-4, what='grad w')
    assert float(l2[..., 2].abs().max()) == 0.0


@pytest.mark.parametrize('L,P,Q', [(4, 4, 10), (2, 3, 7), (5, 4, 33), (1, 1, 5)])
def test_msmv_bwd_deterministic_vs_c_oracle(L, P, Q):
    """sbev_msmv_bwd_det: grad_feats by a per-pixel ordered reduction, no floating-point atomics.
    Same oracle and tolerances as the atomic backward; grad_feats is allocated uninitialised (torch.empty_like) because the
    kernel's contract is that every pixel row is written exactly once (no zero fill needed)."""
    hw = [(7, 9), (4, 6), (3, 3), (2, 2), (1, 2)][:L]
    feats, loc, w = _rand_case(2, 6, hw, Q, P, seed=70 + L)
    go = hashrand((2, Q, 64, P), 98, -1, 1)
    gf, gl, gw = c_oracle.bwd(go, feats, loc, w)
    args = (go.to(dev()), [f.to(dev()) for f in feats], loc.to(dev()), w.to(dev()))
    f2, l2, w2 = _ops().msmv_backward(*args, deterministic=True)
    for a, b in zip(f2, gf):
        _close(a, b, atol=1e-4, what='deterministic grad feats')
    _close(l2, gl, atol=1e-3, what='grad loc')
    _close(w2, gw, atol=1e-4, what='grad w')
    f3, l3, w3 = _ops().msmv_backward(*args, deterministic=True)
    assert all(torch.equal(a, b) for a, b in zip(f2, f3)) and torch.equal(l2, l3) and torch.equal(w2, w3)


def test_msmv_bwd_deterministic_long_segments_and_full_size():
    """(a) every point of a query block lands on the same few pixels -> segments far longer than a half-warp (the
    strided selection path); (b) r50-T1 size: bit-identical between runs, equal to the atomic kernel within fp32
    summation-order noise, and grad_feats sums to the same total (conservation: sum of taps of a point = its weight)."""
    ops = _ops()
    hw = [(5, 6), (3, 3)]
    feats = [hashrand((1, 2, h, w, 64), 3 + i, -1, 1).to(dev()) for i, (h, w) in enumerate(hw)]
    loc = torch.zeros(1, 300, 4, 3)
    loc[..., 0] = 0.41 + 0.001 * torch.arange(4).float()          # all inside the same cell of both levels
    loc[..., 1] = 0.37
    loc[:, 150:, :, 2] = 1.0                                      # half of the queries look at view 1
    w = torch.softmax(hashrand((1, 300, 4, 2), 9, -1, 1), -1)
    go = hashrand((1, 300, 64, 4), 10, -1, 1)
    want = c_oracle.bwd(go, [f.cpu() for f in feats], loc, w)[0]
    got = ops.msmv_backward(go.to(dev()), feats, loc.to(dev()), w.to(dev()), deterministic=True)[0]
    for a, b in zip(got, want):
        _close(a, b, rtol=1e-4, atol=2e-3, what='long segments (600 contributions per pixel)')
    again = ops.msmv_backward(go.to(dev()), feats, loc.to(dev()), w.to(dev()), deterministic=True)[0]
    assert all(torch.equal(a, b) for a, b in zip(got, again))

    hw = [(64, 176), (32, 88), (16, 44), (8, 22)]
    torch.manual_seed(1)
    feats = [torch.randn(4, 6, h, w_, 64, device=dev()) for h, w_ in hw]
    loc = torch.rand(4, 900, 4, 3, device=dev()) * 1.1 - 0.05
    loc[..., 2] = torch.randint(0, 6, (4, 900, 4), device=dev()).float() / 5
    w = torch.softmax(torch.randn(4, 900, 4, 4, device=dev()), -1)
    go = torch.randn(4, 900, 64, 4, device=dev())
    d1 = ops.msmv_backward(go, feats, loc, w, deterministic=True)
    d2 = ops.msmv_backward(go, feats, loc, w, deterministic=True)
    at = ops.msmv_backward(go, feats, loc, w)
    for a, b, c in zip(d1[0], d2[0], at[0]):
        assert torch.equal(a, b), 'deterministic backward differs between two runs'
        _close(a, c, rtol=1e-4, atol=1e-4, what='deterministic vs atomic grad feats')
    assert torch.equal(d1[1], at[1]) and torch.equal(d1[2], at[2])          # grad_loc / grad_w: same kernel code
    empty = ops.msmv_backward(go[:, :0], feats, loc[:, :0], w[:, :0], deterministic=True)
    assert all(float(g.abs().max()) == 0.0 for g in empty[0])


def test_autograd_deterministic_mode():
    """torch.use_deterministic_algorithms(True) routes the autograd Functions to the deterministic backward."""
    from sparsebev_b200 import wrapper
    hw = [(6, 8), (3, 4), (2, 2), (1, 2)]
    feats, loc, w = _rand_case(2, 6, hw, 50, 4, seed=5)
    grads = []
    prev = torch.are_deterministic_algorithms_enabled()
    try:
        torch.use_deterministic_algorithms(True)
        for _ in range(2):
            f = [x.to(dev()).requires_grad_() for x in feats]
            out = wrapper.msmv_sampling(f, loc.to(dev()), w.to(dev()))
            (out * out).sum().backward()
            grads.append([x.grad.clone() for x in f])
    finally:
        torch.use_deterministic_algorithms(prev)
    assert all(torch.equal(a, b) for a, b in zip(*grads))


def test_autograd_function_surface():
    from sparsebev_b200 import wrapper
    hw = [(6, 8), (3, 4), (2, 2), (1, 2)]
    feats, loc, w = _rand_case(2, 6, hw, 6, 4, seed=3)
    feats = [f.to(dev()).requires_grad_() for f in feats]
    loc, w = loc.to(dev()).requires_grad_(), w.to(dev()).requires_grad_()
    out = wrapper.msmv_sampling(feats, loc, w)
    assert out.shape == (2, 6, 64, 4)
    out.sum().backward()
    assert all(f.grad is not None for f in feats) and loc.grad is not None and w.grad is not None
    out2 = wrapper.MSMVSamplingC2345.apply(*[f.detach() for f in feats], loc.detach(), w.detach())
    assert torch.equal(out2, out.detach())
    with pytest.raises(RuntimeError, match='CUDA'):
        wrapper.msmv_sampling([f.detach().cpu() for f in feats], loc.detach().cpu(), w.detach().cpu())


def test_pybind_module_stand_in():
    """`_msmv_sampling_cuda` surface (msmv_sampling.cpp:362-369): forward value and the 6-/7-element backward list."""
    from sparsebev_b200 import _msmv_sampling_cuda as ext, ops
    for hw in ([(6, 8), (3, 4), (2, 2), (1, 2)], [(6, 8), (3, 4), (2, 2), (2, 3), (1, 2)]):
        feats, loc, w = _rand_case(2, 6, hw, 5, 4, seed=11)
        feats, loc, w = [f.to(dev()) for f in feats], loc.to(dev()), w.to(dev())
        fwd = ext._ms_deform_attn_cuda_c2345_forward if len(hw) == 4 else ext._ms_deform_attn_cuda_c23456_forward
        bwd = ext._ms_deform_attn_cuda_c2345_backward if len(hw) == 4 else ext._ms_deform_attn_cuda_c23456_backward
        out = fwd(*feats, loc, w)
        assert torch.equal(out, ops.msmv_forward(feats, loc, w))
        go = torch.randn_like(out)
        grads = bwd(go, *feats, loc, w)
        assert isinstance(grads, list) and len(grads) == len(hw) + 2
        gf, gl, gw = ops.msmv_backward(go, feats, loc, w)
        assert torch.allclose(grads[-1], gw) and torch.allclose(grads[-2], gl) and grads[0].shape == feats[0].shape
        assert float(grads[-2][..., 2].abs().max()) == 0.0          # view coordinate gets no gradient (backward.cu)


def test_against_reference_cuda_kernel():
    """O3 of SURVEY 8(c): our op vs the UNMODIFIED reference kernel compiled for sm_100a (oracle/_ref)."""
    ref = _ref_cuda()
    if ref is None:
        pytest.skip('oracle/_ref/_msmv_sampling_cuda.so not present')
    ops = _ops()
    for L in (4, 5):
        hw = [(64, 176), (32, 88), (16, 44), (8, 22), (4, 11)][:L]
        torch.manual_seed(L)
        feats = [torch.randn(4, 6, h, w, 64, device=dev()) for h, w in hw]
        loc = torch.rand(4, 300, 4, 3, device=dev()) * 1.2 - 0.1
        loc[..., 2] = torch.randint(0, 6, (4, 300, 4), device=dev()).float() / 5
        w = torch.softmax(torch.randn(4, 300, 4, L, device=dev()), -1)
        fwd = getattr(ref, '_ms_deform_attn_cuda_c2345_forward' if L == 4 else '_ms_deform_attn_cuda_c23456_forward')
        bwd = getattr(ref, '_ms_deform_attn_cuda_c2345_backward' if L == 4 else '_ms_deform_attn_cuda_c23456_backward')
        want = fwd(*feats, loc, w)
        got = ops.msmv_forward(feats, loc, w)
        _close(got, want, rtol=1e-5, atol=1e-6, what='fwd vs reference CUDA kernel L=%d' % L)
        go = torch.randn_like(want)
        rg = bwd(go, *feats, loc, w)
        gf, gl, gw = ops.msmv_backward(go, feats, loc, w)
        for a, b in zip(gf, rg[:L]):
            _close(a, b, atol=1e-4, what='grad feats vs reference CUDA')
        _close(gl, rg[L], atol=2e-3, what='grad loc vs reference CUDA')
        _close(gw, rg[L + 1], atol=1e-4, what='grad w vs reference CUDA')


# ---------------------------------------------------------------------------- fused sampling_4d
def _sampling_inputs(golden_dir):
    g = np.load(os.path.join(golden_dir, 'sampling4d.npz'))
    B, Q, T, G, P, L, C, ih, iw = [int(v) for v in g['dims']]
    feats_cf = [hashrand((B * T * G, C, 6, int(h), int(w_)), int(s), -1, 1) for (h, w_), s in zip(g['hw'], g['feat_seeds'])]
    return g, (B, Q, T, G, P, L, C, ih, iw), feats_cf


def test_fused_sampling_matches_golden_reference(golden_dir):
    g, (B, Q, T, G, P, L, C, ih, iw), feats_cf = _sampling_inputs(golden_dir)
    ops = _ops()
    feats = [f.permute(0, 2, 3, 4, 1).contiguous().to(dev()) for f in feats_cf]            # [BTG,N,H,W,C]
    pts = torch.from_numpy(g['points']).to(dev())                                          # [B,Q,GP,3] un-warped
    vel = torch.from_numpy(g['query_bbox'][..., 8:10]).contiguous().to(dev())
    td = torch.from_numpy(g['time_diff']).to(dev())
    l2i = torch.from_numpy(g['lidar2img']).to(dev())
    sw = torch.from_numpy(g['scale_weights'][:, :, :, 0]).contiguous().to(dev())           # [B,Q,G,P,L]
    out, loc = ops.sampling4d_fused(feats, pts, vel, td, l2i, sw, ih, iw, num_frames=T, return_loc=True)
    # the reference's own loc (view choice identical, uv to fp32 round-off of its 4x4 matmul)
    assert torch.equal(loc[..., 2].cpu(), torch.from_numpy(g['loc'][..., 2])), 'view selection differs from reference'
    _close(loc[..., :2], torch.from_numpy(g['loc'][..., :2]), rtol=1e-5, atol=1e-5, what='loc vs reference sampling_4d')
    _close(out, torch.from_numpy(g['out']), rtol=1e-4, atol=1e-4, what='fused sampling vs reference sampling_4d')
    # drop-in sampling_4d signature (warped points per frame)
    from sparsebev_b200 import sampling
    out2 = sampling.sampling_4d(torch.from_numpy(g['points6']).to(dev()), feats, torch.from_numpy(g['scale_weights']).to(dev()),
                                l2i, ih, iw)
    _close(out2, torch.from_numpy(g['out']), rtol=1e-4, atol=1e-4, what='sampling_4d drop-in vs reference')
    pts2 = sampling.make_sample_points(torch.from_numpy(g['query_bbox']).to(dev()), torch.from_numpy(g['offset']).to(dev()),
                                       g['pc_range'].tolist())
    _close(pts2, torch.from_numpy(g['points']), rtol=1e-5, atol=1e-5, what='make_sample_points vs reference')


@pytest.mark.parametrize('variant', [0, 1, 2])
@pytest.mark.parametrize('T', [1, 2, 8])
def test_fused_sampling_loc_bit_exact_vs_oracle_and_layouts(T, variant, option):
    option('gather_variant', variant)
    """Given identical points, the fused kernel's (u, v, view) must equal the oracle's fixed-order fp32
    projection BIT FOR BIT, and the two feature layouts must give identical samples."""
    from sparsebev_b200 import synthetic as S
    ops = _ops()
    cfg = S.layer_cfg('tiny', T)
    B, Q, G, P, L = 2, 36, 4, 4, cfg['num_levels']
    l2i, stamps = S.camera_rig(T, cfg['image_h'], cfg['image_w'])
    l2i = l2i[None].repeat(B, 1, 1, 1).contiguous()
    td = R.time_diff_from_timestamps([stamps] * B)
    qb = S.init_query_bbox(Q, seed=1)[None].repeat(B, 1, 1)
    qb[..., 8:10] = hashrand((B, Q, 2), 3, -1, 1)
    pts = R.make_sample_points(qb, hashrand((B, Q, G * P, 3), 4, -0.5, 0.5), cfg['pc_range'])
    sw = torch.softmax(hashrand((B, Q, G, P, L), 5, -2, 2), -1)
    feats_nchw = S.make_feats('tiny', T, batch=B, seed=2)
    grouped = R.regroup_feats(feats_nchw, channel_last=True)
    pts6 = pts.reshape(B, Q, 1, G, P, 3).expand(B, Q, T, G, P, 3)
    shift = qb[..., 8:10][:, :, None, :] * td[:, None, :, None]
    pts6 = torch.cat([pts6[..., 0:2] - shift[:, :, :, None, None, :], pts6[..., 2:3]], dim=-1)
    sw6 = sw[:, :, :, None].expand(B, Q, G, T, P, L)
    want_loc, _ = R.sampling_4d(pts6, None, sw6, l2i, cfg['image_h'], cfg['image_w'], return_loc=True)
    want = R.sampling_4d(pts6, grouped, sw6, l2i, cfg['image_h'], cfg['image_w'], op=R.msmv_sampling_kernel_semantics)
    args = (pts.to(dev()), qb[..., 8:10].contiguous().to(dev()), td.to(dev()), l2i.to(dev()), sw.to(dev()),
            cfg['image_h'], cfg['image_w'])
    out_g, loc = ops.sampling4d_fused([f.to(dev()) for f in grouped], *args, num_frames=T, return_loc=True)
    assert torch.equal(loc.cpu(), want_loc), 'projected sample locations are not bit-exact'
    _close(out_g, want, what='fused (grouped layout) vs oracle')
    nhwc = [f.permute(0, 1, 3, 4, 2).contiguous().to(dev()) for f in feats_nchw]
    out_n = ops.sampling4d_fused(nhwc, *args, num_frames=T, layout='nhwc')
    assert torch.equal(out_n, out_g), 'nhwc and grouped layouts disagree'


# ------------------------------------------------------------------------------- small dense ops
@pytest.mark.parametrize('M,K,N,ln,relu,res', [(900, 256, 256, True, True, 'post'), (37, 3, 256, True, True, None),
                                              (900, 256, 768, False, False, None), (123, 512, 256, True, False, 'pre'),
                                              (900, 256, 10, False, False, None), (64, 256, 48, False, True, None),
                                              (5, 256, 512, False, True, None)])
def test_dense_vs_torch(M, K, N, ln, relu, res):
    ops = _ops()
    torch.manual_seed(M + K + N)
    x, W, b = torch.randn(M, K), torch.randn(N, K) * 0.1, torch.randn(N) * 0.1
    lw, lb, r = 1 + 0.1 * torch.randn(N), 0.1 * torch.randn(N), torch.randn(M, N)
    y = torch.nn.functional.linear(x, W, b)
    if res == 'pre':
        y = y + r
    if ln:
        y = torch.nn.functional.layer_norm(y, (N,), lw, lb)
    if relu:
        y = torch.relu(y)
    if res == 'post':
        y = y + r
    cache = ops.DenseWeight()
    wt, ldw = cache.get(W.to(dev()))
    got = ops.dense(x.to(dev()), wt, ldw, N, bias=b.to(dev()), ln_w=lw.to(dev()) if ln else None, ln_b=lb.to(dev()) if ln else None,
                    residual=r.to(dev()) if res else None, relu=relu, res_pre_ln=res == 'pre')
    _close(got, y, rtol=1e-4, atol=2e-5, what='dense')


@pytest.mark.parametrize('impl', [0, 1, 2, 4, 8, 12, 14])
def test_dense_chain_vs_torch(impl, option):
    """5-layer chain (FFN + norm3 -> cls branch) and a 3-layer chain with the refine epilogue, intermediate outputs stored.
    impl 0 = tensor-core chain (mma.sync bf16x3, TMA-streamed weights), impl 1 = fp32 FFMA chain."""
    # 0 = tensor-core chain (default), 2 / 4 / 8 = tensor-core chain with that many CTAs per TMA-multicast cluster, 1 = fp32 FFMA
    # 12 / 14 = N-split cluster chain (2 / 4 CTAs per cluster split every layer's output features, DSMEM exchange)
    option('dense_impl', 1 if impl == 1 else 0)
    option('dense_cluster', impl if 2 <= impl <= 8 else 0)
    option('dense_nsplit', impl - 10 if impl >= 10 else 0)
    atol = 2e-5 if impl == 1 else 1e-4
    ops = _ops()
    torch.manual_seed(1)
    M, D = 901, 256
    x = torch.randn(M, D)
    lin = [torch.nn.Linear(256, 512), torch.nn.Linear(512, 256), torch.nn.Linear(256, 256), torch.nn.Linear(256, 256), torch.nn.Linear(256, 10)]
    lns = [None, torch.nn.LayerNorm(256), torch.nn.LayerNorm(256), torch.nn.LayerNorm(256), None]
    for ln in lns:
        if ln is not None:
            torch.nn.init.normal_(ln.weight, 1, 0.1); torch.nn.init.normal_(ln.bias, 0, 0.1)
    with torch.no_grad():
        h = torch.relu(lin[0](x))
        q4 = lns[1](x + lin[1](h))
        c = torch.relu(lns[2](lin[2](q4)))
        c = torch.relu(lns[3](lin[3](c)))
        want_cls = lin[4](c)
        want_a, want_b = lin[2](q4), lin[4](q4)
        want_delta = lin[4](q4)
    caches = [ops.DenseWeight() for _ in lin]
    xd = x.to(dev())
    q4d, clsd = torch.empty(M, 256, device=dev()), torch.empty(M, 10, device=dev())
    import copy
    mods = [copy.deepcopy(m).to(dev()) for m in lin]
    lnd = [None if l is None else copy.deepcopy(l).to(dev()) for l in lns]

    def entry(i, **kw):
        wt, ldw, bias = caches[i].get_with_bias([mods[i].weight], [mods[i].bias])
        return ops.chain_layer(wt, ldw, mods[i].in_features, mods[i].out_features, bias=bias, ln=lnd[i],
                               w_hi=caches[i].w_hi, w_lo=caches[i].w_lo, kpad=caches[i].kpad, **kw)
    ops.dense_chain(xd, D, M, [entry(0, relu=True), entry(1, residual=xd, res_pre_ln=True, y=q4d), entry(2, relu=True),
                               entry(3, relu=True), entry(4, y=clsd)])
    _close(q4d, q4, rtol=1e-4, atol=atol, what='chain intermediate (ffn + norm3)')
    _close(clsd, want_cls, rtol=1e-4, atol=atol, what='chain final (cls)')
    # concatenated Linear (two heads sharing the input) == the two separate Linears
    cat = ops.DenseWeight()
    wt, ldw, bias = cat.get_with_bias([mods[2].weight, mods[4].weight], [mods[2].bias, mods[4].bias])
    both = torch.empty(M, 266, device=dev())
    ops.dense_chain(q4d, D, M, [ops.chain_layer(wt, ldw, 256, 266, bias=bias, y=both, w_hi=cat.w_hi, w_lo=cat.w_lo, kpad=cat.kpad)])
    _close(both[:, :256], want_a, rtol=1e-4, atol=atol, what='concat head A')
    _close(both[:, 256:], want_b, rtol=1e-4, atol=atol, what='concat head B')
    # refine epilogue
    qb = R.init_query_bbox(961, seed=2)[:M][None].contiguous()
    td = torch.tensor([[0.0, 0.5, 1.0]])
    box = torch.empty(M, 10, device=dev())
    ops.dense_chain(q4d, D, M, [entry(4, refine=True, y=box)], refine_proposal=qb.to(dev()), refine_time_diff=td.to(dev()), refine_Q=M, refine_T=3)
    wb = R.refine_bbox(qb, want_delta[None])
    wb = torch.cat([wb[..., :8], wb[..., 8:] / 0.5], -1)
    _close(box, wb[0], rtol=1e-4, atol=atol, what='refine epilogue')
    # K = 3 first layer (position encoder) and a wide N = 776 single layer
    torch.manual_seed(5)
    l3, l776 = torch.nn.Linear(3, 256), torch.nn.Linear(256, 776)
    x10 = torch.randn(M, 10)
    with torch.no_grad():
        w3, w776 = torch.relu(l3(x10[:, :3])), l776(q4)
    c3, c776 = ops.DenseWeight(), ops.DenseWeight()
    y3, y776 = torch.empty(M, 256, device=dev()), torch.empty(M, 776, device=dev())
    l3d, l776d = copy.deepcopy(l3).to(dev()), copy.deepcopy(l776).to(dev())
    wt, ldw, bias = c3.get_with_bias([l3d.weight], [l3d.bias])
    ops.dense_chain(x10.to(dev()), 10, M, [ops.chain_layer(wt, ldw, 3, 256, bias=bias, relu=True, y=y3, w_hi=c3.w_hi, w_lo=c3.w_lo, kpad=c3.kpad)])
    _close(y3, w3, rtol=1e-4, atol=atol, what='K=3 layer')
    wt, ldw, bias = c776.get_with_bias([l776d.weight], [l776d.bias])
    ops.dense_chain(q4d, D, M, [ops.chain_layer(wt, ldw, 256, 776, bias=bias, y=y776, w_hi=c776.w_hi, w_lo=c776.w_lo, kpad=c776.kpad)])
    _close(y776, w776, rtol=1e-4, atol=atol, what='N=776 layer')


@pytest.mark.parametrize('impl,K0,nsplit,ln', [(0, 256, 18, True), (0, 128, 5, True), (0, 256, 1, False), (1, 256, 7, True), (0, 192, 3, True),
                                               (12, 256, 18, True), (14, 256, 18, True), (14, 128, 5, True), (12, 256, 1, False)])
def test_dense_chain_with_fused_splitk_reduce(impl, K0, nsplit, ln, option):
    """sbev_dense_chain_reduce_fwd: input rows = LN(sum_z partial + bias + residual), FFN-like chain behind it (impl 1 and
    K0 = 192 take the documented two-launch fallback)."""
    option('dense_impl', impl if impl < 10 else 0)
    option('dense_nsplit', impl - 10 if impl >= 10 else 0)
    ops = _ops()
    torch.manual_seed(5)
    M = 901
    part, bias, res = torch.randn(nsplit, M, K0), torch.randn(K0), torch.randn(M, K0)
    norm_in = torch.nn.LayerNorm(K0)
    torch.nn.init.normal_(norm_in.weight, 1, 0.1); torch.nn.init.normal_(norm_in.bias, 0, 0.1)
    lin = [torch.nn.Linear(K0, 512), torch.nn.Linear(512, K0)]
    norm_out = torch.nn.LayerNorm(K0)
    with torch.no_grad():
        x_in = part.sum(0) + bias + res
        if ln:
            x_in = norm_in(x_in)
        want = norm_out(x_in + lin[1](torch.relu(lin[0](x_in))))
    import copy
    mods = [copy.deepcopy(m).to(dev()) for m in lin]
    caches = [ops.DenseWeight() for _ in lin]
    nin, nout = copy.deepcopy(norm_in).to(dev()), copy.deepcopy(norm_out).to(dev())
    x_out, y = torch.empty(M, K0, device=dev()), torch.empty(M, K0, device=dev())

    def entry(i, **kw):
        wt, ldw, b = caches[i].get_with_bias([mods[i].weight], [mods[i].bias])
        return ops.chain_layer(wt, ldw, mods[i].in_features, mods[i].out_features, bias=b, w_hi=caches[i].w_hi, w_lo=caches[i].w_lo,
                               kpad=caches[i].kpad, **kw)
    ops.dense_chain_reduce(part.to(dev()), bias.to(dev()), res.to(dev()), nin.weight if ln else None, nin.bias if ln else None, x_out,
                           [entry(0, relu=True), entry(1, ln=nout, residual=x_out, res_pre_ln=True, y=y)])
    _close(x_out, x_in, rtol=1e-4, atol=1e-5, what='fused split-K reduce + LN (chain input)')
    _close(y, want, rtol=1e-4, atol=2e-5 if impl == 1 else 1e-4, what='chain behind the fused reduce')


def test_dense_chain_packed_weight_stream_is_bit_identical(option):
    """sbev_dense_layer.W_pack (pre-tiled, pre-swizzled weights streamed by one 32 KB bulk copy per stage) puts exactly the
    bytes into shared memory that the tensor-map path does: outputs are bit-identical, ragged N (10, 776) and K (3) included."""
    ops = _ops()
    torch.manual_seed(7)
    M = 333
    lin = [torch.nn.Linear(3, 256), torch.nn.Linear(256, 256), torch.nn.Linear(256, 776)]
    lin2 = [torch.nn.Linear(256, 512), torch.nn.Linear(512, 256), torch.nn.Linear(256, 10)]
    for chain, K0 in ((lin, 3), (lin2, 256)):
        mods = [m.to(dev()) for m in chain]
        caches = [ops.DenseWeight() for _ in mods]
        x = torch.randn(M, K0).to(dev())
        outs = {}
        for pack in (1, 0):
            option('dense_pack', pack)
            ys = [torch.empty(M, m.out_features, device=dev()) for m in mods]
            entries = []
            for i, m in enumerate(mods):
                wt, ldw, bias = caches[i].get_with_bias([m.weight], [m.bias])
                assert caches[i].w_pack is not None and caches[i].w_pack.shape[:3] == ((m.out_features + 127) // 128, caches[i].kpad // 64, 2)
                entries.append(ops.chain_layer(wt, ldw, m.in_features, m.out_features, bias=bias, relu=i + 1 < len(mods), y=ys[i],
                                               w_hi=caches[i].w_hi, w_lo=caches[i].w_lo, kpad=caches[i].kpad, w_pack=caches[i].w_pack))
            ops.dense_chain(x, K0, M, entries)
            torch.cuda.synchronize()
            outs[pack] = ys
        for a, b in zip(outs[1], outs[0]):
            assert torch.equal(a, b)
        with torch.no_grad():
            h = x
            for i, m in enumerate(mods):
                h = m(h)
                if i + 1 < len(mods):
                    h = torch.relu(h)
        _close(outs[1][-1], h.cpu(), rtol=1e-4, atol=1e-4, what='packed-stream chain vs torch')


@pytest.mark.parametrize('fuse', [1, 0])
def test_dense_chain_points_equals_chain_then_sample_points(fuse, option):
    """sbev_dense_chain_points_fwd: the sample points / scale weights that leave the chain's epilogue are bit-identical to
    running the chain and then sbev_sample_points_fwd on its output (fuse=0 exercises the documented two-launch form)."""
    from sparsebev_b200 import _lib
    ops = _ops()
    prev = _lib.get_option('dense_fuse_points')
    _lib.set_option('dense_fuse_points', fuse)
    try:
        torch.manual_seed(3)
        M, D, GP, L = 901, 256, 16, 4
        x = torch.randn(M, D)
        lin = [torch.nn.Linear(256, 256), torch.nn.Linear(256, GP * 3 + GP * L)]
        ln = torch.nn.LayerNorm(256)
        qb = R.init_query_bbox(961, seed=2)[:M][None].contiguous()
        import copy
        mods = [copy.deepcopy(m).to(dev()) for m in lin]
        lnd = copy.deepcopy(ln).to(dev())
        caches = [ops.DenseWeight() for _ in lin]

        def entry(i, **kw):
            wt, ldw, bias = caches[i].get_with_bias([mods[i].weight], [mods[i].bias])
            return ops.chain_layer(wt, ldw, mods[i].in_features, mods[i].out_features, bias=bias, w_hi=caches[i].w_hi, w_lo=caches[i].w_lo,
                                   kpad=caches[i].kpad, **kw)
        xd = x.to(dev())
        q2, heads = torch.empty(M, 256, device=dev()), torch.empty(M, GP * 3 + GP * L, device=dev())
        pc = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
        pts, sw = ops.dense_chain_points(xd, D, M, [entry(0, ln=lnd, residual=xd, res_pre_ln=True, y=q2), entry(1, y=heads)], qb.to(dev()),
                                         pc, GP, L, 0, GP * 3)
        q2b, headsb = torch.empty_like(q2), torch.empty_like(heads)
        ops.dense_chain(xd, D, M, [entry(0, ln=lnd, residual=xd, res_pre_ln=True, y=q2b), entry(1, y=headsb)])
        pts2, sw2 = ops.sample_points(qb.to(dev()), headsb, headsb[:, GP * 3:], pc, L, num_points_total=GP, ld_off=headsb.shape[1], ld_log=headsb.shape[1])
        torch.cuda.synchronize()
        assert torch.equal(heads, headsb) and torch.equal(q2, q2b)
        assert torch.equal(pts, pts2) and torch.equal(sw, sw2)
        want_pts = R.make_sample_points(qb, headsb[:, :GP * 3].cpu().reshape(1, M, GP, 3), pc)
        _close(pts, want_pts, rtol=1e-5, atol=1e-4, what='fused sample points vs oracle')
    finally:
        _lib.set_option('dense_fuse_points', prev)


def test_sample_points_and_refine_vs_oracle():
    ops = _ops()
    pc = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    qb = R.init_query_bbox(900, seed=2)[None].repeat(2, 1, 1)
    qb[..., 8:10] = hashrand((2, 900, 2), 1, -1, 1)
    off = hashrand((2, 900, 16, 3), 2, -0.5, 0.5)
    logits = hashrand((2, 900, 16, 4), 3, -3, 3)
    pts, sw = ops.sample_points(qb.to(dev()), off.reshape(2, 900, 48).to(dev()), logits.reshape(2, 900, 64).to(dev()), pc, 4)
    _close(pts, R.make_sample_points(qb, off, pc), rtol=1e-5, atol=2e-5, what='sample points')
    _close(sw, torch.softmax(logits, -1), rtol=1e-5, atol=1e-6, what='scale weights')
    delta = hashrand((2, 900, 10), 4, -2, 2)
    td = torch.tensor([[0.0, 0.5, 1.0], [0.0, 1e-6, 1.0]])
    want = R.refine_bbox(qb, delta)
    t = td.clone()
    t[t < 1e-5] = 1.0
    want = torch.cat([want[..., :8], want[..., 8:] / t[:, 1:2, None]], -1)
    _close(ops.refine_bbox(qb.to(dev()), delta.to(dev()), td.to(dev())), want, rtol=1e-5, atol=1e-6, what='refine_bbox')


def test_sample_points_legacy_rotation(option):
    """Option "legacy_rotation" (checkpoints with version 'v0.17.1', /root/reference/models/utils.py:66-71): rot_mat_T =
    [[c, -s, 0], [s, c, 0], [0, 0, 1]] instead of [[c, s, 0], [-s, c, 0], [0, 0, 1]], i.e. the offsets are rotated by -yaw."""
    ops = _ops()
    pc = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    qb = R.init_query_bbox(900, seed=2)[None].contiguous()
    off = hashrand((1, 900, 16, 3), 2, -0.5, 0.5)
    logits = hashrand((1, 900, 16, 4), 3, -3, 3)
    dec = R.decode_bbox(qb, pc)
    delta = dec[..., None, 3:6] * off
    yaw = dec[..., 6]
    c, s_ = torch.cos(yaw)[..., None], torch.sin(yaw)[..., None]
    want = torch.stack([delta[..., 0] * c + delta[..., 1] * s_, -delta[..., 0] * s_ + delta[..., 1] * c, delta[..., 2]], -1) + dec[..., None, 0:3]
    option('legacy_rotation', 1)
    pts, _ = ops.sample_points(qb.to(dev()), off.reshape(1, 900, 48).to(dev()), logits.reshape(1, 900, 64).to(dev()), pc, 4)
    _close(pts, want, rtol=1e-5, atol=2e-5, what='sample points, legacy v0.17.1 rotation')
    option('legacy_rotation', 0)
    pts0, _ = ops.sample_points(qb.to(dev()), off.reshape(1, 900, 48).to(dev()), logits.reshape(1, 900, 64).to(dev()), pc, 4)
    _close(pts0, R.make_sample_points(qb, off, pc), rtol=1e-5, atol=2e-5, what='sample points, v1.0.0 rotation')
    assert float((pts.cpu() - pts0.cpu()).abs().max()) > 0.1


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('Q', [900, 70, 64])
def test_sasa_vs_oracle(Q, impl, option):
    option('sasa_impl', impl)            # 0 = mma.sync bf16x3 (default), 1 = fp32 FFMA
    ops = _ops()
    pc = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    B, D, H = 2, 256, 8
    n = int(np.ceil(np.sqrt(Q))) ** 2
    qb = R.init_query_bbox(n, seed=3)[:Q][None].repeat(B, 1, 1)
    qb[1, :, :2] = hashrand((Q, 2), 8, 0, 1)
    qkv = hashrand((B, Q, 3 * D), 5, -1.5, 1.5)
    tau = hashrand((B, Q, H), 6, 0, 2)
    q, k, v = [t.reshape(B, Q, H, 32).transpose(1, 2) for t in qkv.chunk(3, -1)]
    bias = -R.pairwise_centre_dist(qb, pc)[:, None] * tau.permute(0, 2, 1)[..., None]
    want = torch.softmax(q * (1 / np.sqrt(32)) @ k.transpose(-1, -2) + bias, -1) @ v
    want = want.transpose(1, 2).reshape(B, Q, D)
    got = ops.sasa(qkv.to(dev()), qb.to(dev()), tau.to(dev()), pc, H)
    _close(got, want, rtol=1e-4, atol=1e-5, what='sasa core')
    packed = torch.cat([qkv, tau], -1).reshape(B * Q, 3 * D + H).contiguous().to(dev())       # strided (concatenated-Linear) form
    got_s = ops.sasa(packed, qb.to(dev()), packed[:, 3 * D:], pc, H, ld_qkv=3 * D + H, ld_tau=3 * D + H, embed_dims=D)
    assert torch.equal(got_s, got)
    got_v3 = ops.sasa_split(packed, qb.to(dev()), pc, H, D)                  # warp-pipelined kernel on pre-split operands
    _close(got_v3, want, rtol=1e-4, atol=2e-5, what='sasa core (split / v3)')
    mask = torch.zeros(Q, Q, dtype=torch.bool)
    mask[: Q // 2, Q // 2:] = True
    want_m = torch.softmax((q * (1 / np.sqrt(32)) @ k.transpose(-1, -2) + bias).masked_fill(mask, float('-inf')), -1) @ v
    got_m = ops.sasa(qkv.to(dev()), qb.to(dev()), tau.to(dev()), pc, H, dn_mask=mask.to(dev()))
    _close(got_m, want_m.transpose(1, 2).reshape(B, Q, D), rtol=1e-4, atol=1e-5, what='sasa core with dn mask')
    got_m3 = ops.sasa_split(packed, qb.to(dev()), pc, H, D, dn_mask=mask.to(dev()))
    _close(got_m3, want_m.transpose(1, 2).reshape(B, Q, D), rtol=1e-4, atol=2e-5, what='sasa core (split / v3) with dn mask')


@pytest.mark.parametrize('kq', [4, 8, 0])
def test_sasa_query_range_and_key_splits(kq, option):
    """sbev_sasa_split_range_fwd: a query shard [qa, qb) attends to all keys; with the same key-split count its rows are
    BIT-IDENTICAL to the full launch (a query's result does not depend on its tile mates), with 8 key splits per CTA (the
    small-grid variant, chosen automatically for a shard) they agree with the oracle to the same bar."""
    option('sasa_kq', kq)
    ops = _ops()
    pc = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    Q, D, H = 900, 256, 8
    qb = R.init_query_bbox(900, seed=3)[None].contiguous()
    qkv = hashrand((1, Q, 3 * D), 5, -1.5, 1.5)
    tau = hashrand((1, Q, H), 6, 0, 2)
    q, k, v = [t.reshape(1, Q, H, 32).transpose(1, 2) for t in qkv.chunk(3, -1)]
    bias = -R.pairwise_centre_dist(qb, pc)[:, None] * tau.permute(0, 2, 1)[..., None]
    want = (torch.softmax(q * (1 / np.sqrt(32)) @ k.transpose(-1, -2) + bias, -1) @ v).transpose(1, 2).reshape(1, Q, D)
    packed = torch.cat([qkv, tau], -1).reshape(Q, 3 * D + H).contiguous().to(dev())
    full = ops.sasa_split(packed, qb.to(dev()), pc, H, D)
    _close(full, want, rtol=1e-4, atol=2e-5, what='sasa full launch, sasa_kq=%d' % kq)
    for qa, qe in ((0, 113), (339, 452), (791, 900), (450, 900), (37, 38), (5, 5)):
        out = torch.full((1, Q, D), 7.0, device=dev())
        ops.sasa_split(packed, qb.to(dev()), pc, H, D, q_range=(qa, qe), out=out)
        assert bool((out[:, :qa] == 7.0).all()) and bool((out[:, qe:] == 7.0).all())          # rows outside the shard untouched
        _close(out[:, qa:qe], want[:, qa:qe], rtol=1e-4, atol=2e-5, what='sasa shard [%d,%d), sasa_kq=%d' % (qa, qe, kq))
        if kq != 0:
            assert torch.equal(out[:, qa:qe], full[:, qa:qe])


# ------------------------------------------------------------------------- tcgen05 GEMM + mixing
@pytest.mark.parametrize('impl', [0])
@pytest.mark.parametrize('M,N,K,split_k', [(128, 128, 64, 1), (900, 256, 256, 1), (900, 1024, 256, 1), (300, 256, 2048, 8),
                                           (1, 128, 128, 2), (900, 384, 512, 2), (2000, 2560, 128, 1), (900, 256, 4096, 18), (260, 256, 1024, 5)])
def test_gemm_bf16_single_segment(M, N, K, split_k, impl, option):
    option('gemm_impl', impl)            # 1 = persistent double-buffered kernel (default), 0 = one tile per CTA
    ops = _ops()
    torch.manual_seed(M + N + K)
    a = torch.randn(M, K, device=dev()).bfloat16()
    b = torch.randn(N, K, device=dev()).bfloat16()
    bias = torch.randn(N, device=dev())
    got = ops.gemm_bf16_tn([a], [b], M, N, K, bias=bias, split_k=split_k)
    if split_k > 1:
        got = got.sum(0)
    want = a.double() @ b.double().t() + bias.double()
    _close(got, want.float(), rtol=1e-5, atol=1e-4 * np.sqrt(K / 64), what='bf16 gemm (exact products, fp32 accumulate)')


@pytest.mark.parametrize('impl', [0, 1, 2, 3])    # 0 / 2: CTA pairs (cta_group::2) when N % 256 == 0; 1: single CTAs, A-resident for small K / many N; 3: single CTAs
@pytest.mark.parametrize('M,N,K,split_k', [(900, 512, 256, 1), (900, 256, 4096, 8), (900, 384, 256, 1), (900, 8192, 256, 1), (130, 5120, 256, 1)])
def test_gemm_bf16x3_is_fp32_grade(M, N, K, split_k, impl, option):
    option('gemm_impl', impl)
    ops = _ops()
    torch.manual_seed(0)
    a, b = torch.randn(M, K, device=dev()), torch.randn(N, K, device=dev()) * 0.05
    ah, al = ops.split_bf16(a)
    bh, bl = ops.split_bf16(b)
    _close(ah.float() + al.float(), a, rtol=2 ** -15, atol=0, what='bf16 split residual')
    got = ops.gemm_bf16_tn([ah, ah, al], [bh, bl, bh], M, N, K, split_k=split_k)
    if split_k > 1:
        got = got.sum(0)
    want = (a.double() @ b.double().t()).float()
    err = (got - want).abs().max() / want.abs().max()
    assert float(err) < 2e-5, 'bf16x3 relative error %.3e' % float(err)
    single = ops.gemm_bf16_tn([ah], [bh], M, N, K)
    assert float((single - want).abs().max() / want.abs().max()) > float(err) * 10     # the split really matters


def test_mix_kernel_and_full_mixing_vs_reference_golden(golden_dir):
    """AdaptiveMixing end to end (param-gen GEMM -> mix -> out_proj GEMM -> +query) against the golden
    output of the REAL reference class, and the middle stage alone against the oracle."""
    import sparsebev_b200 as sb
    g = np.load(os.path.join(golden_dir, 'mixing.npz'))
    Bm, Qm, G, Pin, C = [int(v) for v in g['dims']]
    s = [int(v) for v in g['seeds']]
    mod = sb.AdaptiveMixing(in_dim=256, in_points=Pin, n_groups=4, out_points=128)
    mod.load_state_dict({
        'parameter_generator.weight': hashrand((G * (C * C + Pin * 128), 256), s[0], -0.04, 0.04),
        'parameter_generator.bias': hashrand((G * (C * C + Pin * 128),), s[1], -0.1, 0.1),
        'out_proj.weight': hashrand((256, G * 128 * C), s[2], -0.02, 0.02),
        'out_proj.bias': hashrand((256,), s[3], -0.05, 0.05)})
    mod = mod.to(dev())
    x = hashrand((Bm, Qm, G, Pin, C), s[4], -2, 2).to(dev())
    q = hashrand((Bm, Qm, 256), s[5], -1.5, 1.5).to(dev())
    out = mod(x, q)
    _close(out, torch.from_numpy(g['out']), rtol=1e-4, atol=1e-4, what='AdaptiveMixing (bf16x3) vs reference')


@pytest.mark.parametrize('impl', [0, 1])
@pytest.mark.parametrize('Pin', [32, 8, 60, 120])
def test_mix_stage_vs_oracle(Pin, impl, option):
    option('mix_impl', impl)             # 0 = mma.sync bf16x3 (default), 1 = fp32 FFMA
    ops = _ops()
    BQ, G, C = 5, 4, 64
    params = hashrand((BQ, G * (C * C + 128 * Pin)), 11 + Pin, -0.3, 0.3)
    x = hashrand((BQ, G, Pin, C), 12 + Pin, -2, 2)
    p3 = params.reshape(BQ, G, -1)
    m = p3[..., :C * C].reshape(BQ, G, C, C)
    sm = p3[..., C * C:].reshape(BQ, G, 128, Pin)
    h = torch.relu(torch.nn.functional.layer_norm(x @ m, (Pin, C)))
    want = torch.relu(torch.nn.functional.layer_norm(sm @ h, (128, C))).reshape(BQ, -1)
    hi, lo, yf = ops.mix(params.to(dev()), x.to(dev()), want_f32=True)
    # fp32 FFMA path: 1e-4 rel / 1e-5 abs.  Tensor-core path (bf16x3 products, operands held as bf16 hi+lo pairs, i.e.
    # ~2^-17 relative each): the outputs are LayerNorm'ed to unit variance, absolute floor 5e-5 (= 5e-5 of the tensor scale).
    atol = 1e-5 if impl == 1 else 5e-5
    _close(yf, want, rtol=1e-4, atol=atol, what='mix stage fp32')
    _close(hi.float() + lo.float(), want, rtol=1e-4, atol=atol, what='mix stage bf16 hi+lo')
    assert float((yf.cpu() - want).abs().max() / want.abs().max()) < 2e-5


@pytest.mark.parametrize('nsplit,M,N,ln', [(16, 900, 256, True), (18, 900, 256, True), (1, 37, 64, True), (33, 5, 1024, True),
                                          (3, 130, 96, True), (5, 64, 256, False)])
def test_reduce_ln_vs_torch(nsplit, M, N, ln):
    ops = _ops()
    torch.manual_seed(0)
    part, bias, res = torch.randn(nsplit, M, N), torch.randn(N), torch.randn(M, N)
    lw, lb = torch.randn(N), torch.randn(N)
    want = part.sum(0) + bias + res
    if ln:
        want = torch.nn.functional.layer_norm(want, (N,), lw, lb)
    got = ops.reduce_ln(part.to(dev()), bias.to(dev()), res.to(dev()), lw.to(dev()) if ln else None, lb.to(dev()) if ln else None)
    _close(got, want, rtol=1e-4, atol=1e-5, what='reduce_ln')


@pytest.mark.parametrize('impl', [0, 2, 4, 5])
def test_gemm_split_output_and_tma_fed_mix(impl, option):
    """GEMM with bf16 (hi, lo) output + the TMA-fed mix kernel == the fp32-parameter path (impl 2: CTA-pair GEMM, 0: single CTAs,
    4: A-resident single CTAs with 128-wide N tiles, 5: CTA pairs with a three-stage ring; M = 300 gives
    an odd number of 128-row tiles, so the last pair's second CTA is entirely out of range)."""
    option('gemm_impl', impl)
    ops = _ops()
    torch.manual_seed(3)
    M, N, K = 300, 4 * 8192, 256                       # 300 queries x 4 groups x (64*64 + 128*32) parameters
    a, b = torch.randn(M, K, device=dev()), torch.randn(N, K, device=dev()) * 0.03
    bias = torch.randn(N, device=dev()) * 0.05
    ah, al = ops.split_bf16(a)
    bh, bl = ops.split_bf16(b)
    want = (a.double() @ b.double().t() + bias.double()).float()
    ch, cl = ops.gemm_bf16_tn_split(ah, al, bh, bl, M, N, K, bias=bias)
    err = ((ch.float() + cl.float()) - want).abs().max() / want.abs().max()
    assert float(err) < 3e-5, 'split-output GEMM relative error %.3e' % float(err)
    x = hashrand((M, 4, 32, 64), 77, -2, 2).to(dev())
    hi_ref, lo_ref, y_ref = ops.mix(want, x, want_f32=True)
    hi, lo, y = ops.mix_presplit(ch, cl, x, want_f32=True)
    _close(y, y_ref, rtol=1e-4, atol=5e-5, what='TMA-fed mix vs fp32-parameter mix')
    _close(hi.float() + lo.float(), y_ref, rtol=1e-4, atol=5e-5, what='TMA-fed mix hi+lo')
