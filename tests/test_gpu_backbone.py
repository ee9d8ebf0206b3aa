"""Backbone conv path on the GPU: the tcgen05 implicit-GEMM convolution, stem, max pool and the ResNet /
FPN mirrors against the plain-PyTorch restatement in oracle/ref_backbone.py (F.conv2d -> BatchNorm2d(eval) -> ReLU, the very
calls the reference's wrappers make), fed the same bf16-rounded operands.

Tolerances: fp32 outputs differ from the oracle only by fp32 summation order (rtol 1e-4, atol 1e-4 on O(1) values);
bf16 outputs may differ by one bf16 ulp of the result (2^-8 relative); whole networks accumulate ulp flips layer by layer:
max |err| / max |ref| < 2e-2 per FPN level."""
import numpy as np
import pytest
import torch

from oracle import ref_backbone as RB

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda:0')


def _rand_bn(c, g):
    return (torch.rand(c, generator=g) + 0.5, 0.1 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g),
            torch.rand(c, generator=g) + 0.5, 1e-5)


def _nhwc_bf16(x):
    return x.permute(0, 2, 3, 1).contiguous().to(dev()).bfloat16()


CASES = [
    # Cin, Cout, k, stride, pad, H, W, bn, bias, relu, residual ('same' | 'up' | None), out_f32
    (64, 64, 1, 1, 0, 24, 40, True, False, True, None, False),
    (64, 64, 3, 1, 1, 17, 23, True, False, True, None, False),        # odd sizes: partial tiles on both axes
    (128, 128, 3, 2, 1, 32, 44, True, False, True, None, False),      # strided 3x3 (four parity maps)
    (128, 128, 3, 2, 1, 31, 45, True, False, True, None, True),       # strided 3x3, odd input
    (256, 512, 1, 2, 0, 16, 22, True, False, False, None, False),     # bottleneck downsample
    (64, 256, 1, 1, 0, 16, 24, True, False, True, 'same', False),     # conv3 + identity + relu
    (512, 256, 1, 1, 0, 8, 22, False, True, False, 'up', False),      # FPN lateral + nearest-upsampled coarser level
    (256, 256, 3, 1, 1, 8, 22, False, True, False, None, True),       # FPN output conv, fp32 NHWC
    (2048, 256, 1, 1, 0, 8, 22, False, True, False, None, False),     # deepest lateral (32 k-blocks)
    (64, 96, 3, 1, 1, 9, 16, True, True, True, None, True),           # Cout % 64 != 0
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: 'c%d_%d_k%d_s%d_%dx%d_%s' % (c[0], c[1], c[2], c[3], c[5], c[6], c[10]))
def test_conv2d_nhwc_vs_torch(case):
    from sparsebev_b200 import ops
    Cin, Cout, k, stride, pad, H, W, use_bn, use_bias, relu, res_kind, out_f32 = case
    g = torch.Generator().manual_seed(Cin * 7 + Cout + k + H)
    Nimg = 3
    x = torch.randn(Nimg, Cin, H, W, generator=g).bfloat16().float()
    w = (torch.randn(Cout, Cin, k, k, generator=g) / np.sqrt(Cin * k * k)).bfloat16().float()
    bias = 0.1 * torch.randn(Cout, generator=g) if use_bias else None
    bn = _rand_bn(Cout, g) if use_bn else None
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    res = None
    if res_kind == 'same':
        res = torch.randn(Nimg, Cout, Ho, Wo, generator=g).bfloat16().float()
    elif res_kind == 'up':
        res = torch.randn(Nimg, Cout, Ho // 2, Wo // 2, generator=g).bfloat16().float()
    want = RB.conv_bn_act(x, w, bias, bn, relu, stride, pad, res, emulate_bf16=True)
    if bn is not None:
        scale = bn[0] / torch.sqrt(bn[3] + bn[4])
        shift = bn[1] - bn[2] * scale + (bias * scale if bias is not None else 0)
    else:
        scale, shift = None, (bias if bias is not None else torch.zeros(Cout))
    got = ops.conv2d_nhwc(_nhwc_bf16(x), w.permute(0, 2, 3, 1).contiguous().to(dev()).bfloat16(), shift.to(dev()).contiguous(),
                          None if scale is None else scale.to(dev()).contiguous(), stride=stride, pad=pad,
                          residual=None if res is None else _nhwc_bf16(res), relu=relu, out_f32=out_f32)
    torch.cuda.synchronize()
    assert got.shape == (Nimg, Ho, Wo, Cout) and got.dtype == (torch.float32 if out_f32 else torch.bfloat16)
    got = got.float().cpu().permute(0, 3, 1, 2)
    err = (got - want).abs()
    tol = (1e-4 + 1e-4 * want.abs()) if out_f32 else (1e-3 + 2.0 ** -7 * want.abs())
    bad = err > tol
    assert not bad.any(), '%d / %d off, max err %.3e at ref max %.3e' % (int(bad.sum()), bad.numel(), float(err.max()), float(want.abs().max()))


def test_stem_maxpool_subsample_cast_vs_torch():
    from sparsebev_b200 import ops
    g = torch.Generator().manual_seed(3)
    img = torch.randn(2, 3, 70, 100, generator=g)
    w = torch.randn(64, 3, 7, 7, generator=g) / np.sqrt(147.0)
    bn = _rand_bn(64, g)
    want = RB.conv_bn_act(img, w, None, bn, True, 2, 3)
    scale = bn[0] / torch.sqrt(bn[3] + bn[4]); shift = bn[1] - bn[2] * scale
    got = ops.stem_conv(img.to(dev()), w.permute(2, 3, 1, 0).contiguous().to(dev()), scale.to(dev()), shift.to(dev()))
    assert got.shape == (2, 35, 50, 64) and got.dtype == torch.bfloat16
    e = (got.float().cpu().permute(0, 3, 1, 2) - want).abs()
    assert bool((e <= 1e-3 + 2.0 ** -7 * want.abs()).all()), float(e.max())
    pooled = ops.maxpool3x3s2_nhwc(got)
    wantp = torch.nn.functional.max_pool2d(got.float().cpu().permute(0, 3, 1, 2), 3, 2, 1)
    assert torch.equal(pooled.float().cpu().permute(0, 3, 1, 2), wantp)                    # max of bf16 values: exact
    x = torch.randn(2, 9, 13, 256, generator=g).to(dev())
    sub = ops.subsample2_nhwc(x)
    assert torch.equal(sub.cpu(), x.cpu()[:, ::2, ::2].contiguous())
    assert torch.equal(ops.cast_bf16(x).cpu(), x.cpu().bfloat16())


def _init_backbone(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.Conv2d):
                fan = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * np.sqrt(1.0 / fan))
                if m.bias is not None:
                    m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
            elif isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_mean.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
                m.running_var.copy_(torch.rand(m.bias.shape, generator=g) * 0.5 + 0.75)


@pytest.mark.parametrize('depth,num_outs', [(50, 4), (50, 5), (101, 5)])
def test_resnet_fpn_vs_oracle_and_zero_copy_into_decoder_layout(depth, num_outs):
    """(50, 4): configs/r50_nuimg_704x256.py:31-45; (101, 5): BASELINE config 4, configs/r101_nuimg_1408x512.py:14-25."""
    import sparsebev_b200 as sb
    from sparsebev_b200 import backbone as BB
    net = BB.ResNet(depth=depth, with_cp=True)
    neck = BB.FPN([256, 512, 1024, 2048], 256, num_outs)
    _init_backbone(net, 1); _init_backbone(neck, 2)
    net.eval(); neck.eval()
    B, TN, H, W = 1, 6, 64, 96
    img = torch.randn(B, TN, 3, H, W, generator=torch.Generator().manual_seed(5))
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    c = RB.resnet_forward(img.reshape(B * TN, 3, H, W), sd, depth, emulate_bf16=True)
    want = RB.fpn_forward(c, {k: v.detach().clone() for k, v in neck.state_dict().items()}, num_outs, emulate_bf16=True)
    net.to(dev()); neck.to(dev())
    feats = BB.extract_img_feat(net, neck, img.to(dev()))
    torch.cuda.synchronize()
    assert len(feats) == num_outs
    for lvl, (f, w_) in enumerate(zip(feats, want)):
        assert f.shape == (B, TN, 256) + tuple(w_.shape[-2:]) and f.dtype == torch.float32
        assert f.permute(0, 1, 3, 4, 2).is_contiguous()                     # channels-last memory: the gather's zero-copy layout
        err = float((f.reshape(B * TN, 256, *w_.shape[-2:]).cpu() - w_).abs().max() / w_.abs().max())
        assert err < (2e-2 if depth == 50 else 3e-2), 'FPN level %d: rel-to-max error %.3e' % (lvl, err)
    # the decoder consumes these tensors without a copy
    dec = sb.SparseBEVTransformer(256, num_frames=1, num_points=4, num_layers=1, num_levels=num_outs, pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]).decoder
    lst = list(feats)
    ptrs = [f.data_ptr() for f in lst]
    dec.prepare_feats(lst)
    assert dec.decoder_layer.sampling.feat_layout == 'nhwc' and [f.data_ptr() for f in lst] == ptrs


def test_conv2d_wrapper_and_vovnet_style_sequence_drop_in():
    from sparsebev_b200 import backbone as BB
    g = torch.Generator().manual_seed(11)
    conv = BB.Conv2d(64, 128, 3, stride=2, padding=1, bias=False, norm=torch.nn.BatchNorm2d(128), activation=torch.nn.ReLU())
    seq = BB.ConvBNReLUSequence(BB.conv3x3(64, 128, 'OSA2_1', 0) + BB.conv1x1(128, 64, 'OSA2_1', 'concat'))
    _init_backbone(conv, 3); _init_backbone(seq, 4)
    conv.eval(); seq.eval()
    x = torch.randn(2, 64, 20, 28, generator=g).bfloat16().float()
    with torch.no_grad():
        want = torch.relu(conv.norm(torch.nn.functional.conv2d(x, conv.weight.bfloat16().float(), None, 2, 1)))
        h = x
        for i in range(0, 6, 3):
            c_, n_ = seq[i], seq[i + 1]
            h = torch.relu(n_(torch.nn.functional.conv2d(h.bfloat16().float(), c_.weight.bfloat16().float(), None, c_.stride, c_.padding)))
        want_seq = h
    conv.to(dev()); seq.to(dev())
    got = conv(x.to(dev()))
    assert got.shape == want.shape and got.dtype == torch.float32
    assert torch.allclose(got.cpu(), want, rtol=1e-4, atol=1e-4), float((got.cpu() - want).abs().max())
    got_seq = seq(x.to(dev()))
    assert float((got_seq.cpu() - want_seq).abs().max() / want_seq.abs().max()) < 1e-2
    with pytest.raises(NotImplementedError):
        BB.Conv2d(64, 64, 3, groups=64).to(dev())(x.to(dev()))             # depthwise: no kernel, and no silent cuDNN fallback


def test_maxpool_ceil_mode_and_ese_vs_torch():
    """VoVNet's non-conv pieces: MaxPool2d(3, 2, ceil_mode=True) incl. odd sizes (window overhang), and the eSE block."""
    import torch.nn.functional as F
    from sparsebev_b200 import ops
    g = torch.Generator().manual_seed(0)
    for (H, W, pad, ceil) in ((25, 41, 0, True), (12, 20, 0, True), (6, 10, 0, True), (24, 40, 1, False), (5, 5, 0, True)):
        x = torch.randn(2, 64, H, W, generator=g).cuda()
        xb = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
        want = F.max_pool2d(xb.float().permute(0, 3, 1, 2), 3, 2, padding=pad, ceil_mode=ceil)
        got = ops.maxpool3x3s2_ex_nhwc(xb, pad=pad, ceil_mode=ceil)
        assert got.shape == (2, want.shape[2], want.shape[3], 64)
        assert torch.equal(got.float().permute(0, 3, 1, 2), want)
    C = 256
    x = torch.randn(3, 13, 21, C, generator=g).cuda().to(torch.bfloat16)
    idt = torch.randn(3, 13, 21, C, generator=g).cuda().to(torch.bfloat16)
    w, b = (torch.randn(C, C, generator=g) * 0.2).cuda(), torch.randn(C, generator=g).cuda()
    gate = F.relu6(x.float().mean((1, 2)) @ w.t() + b + 3.0) / 6.0
    for identity in (None, idt):
        want = x.float() * gate[:, None, None, :] + (0 if identity is None else identity.float())
        got = ops.ese_nhwc(x, w, b, identity)
        assert torch.allclose(got.float(), want, rtol=1e-2, atol=1e-2)          # bf16 output rounding


def test_vovnet99_vs_real_reference_golden(golden_dir):
    """BASELINE config 5's image backbone: the V-99-eSE mirror on our kernels (bf16 operands, fp32 accumulate) against the
    stage outputs of the REAL reference class run in fp32 on the CPU (tests/golden/vovnet.npz), same seeded weights."""
    import os
    import numpy as np
    from oracle.gen_golden_vovnet import seeded_init, test_image
    from sparsebev_b200 import backbone as BB
    g = np.load(os.path.join(golden_dir, 'vovnet.npz'))
    feats = ['stage2', 'stage3', 'stage4', 'stage5']
    net = seeded_init(BB.VoVNet('V-99-eSE', out_features=feats), seed=3).cuda().eval()
    out = net.forward_nhwc(test_image().cuda())
    for k in feats:
        want = torch.from_numpy(g[k])
        got = out[k].float().permute(0, 3, 1, 2).cpu()
        assert got.shape == want.shape
        rel = float((got - want).abs().max() / want.abs().max())
        rms = float((got - want).pow(2).mean().sqrt() / want.pow(2).mean().sqrt())
        assert rel < 6e-2 and rms < 2e-2, '%s: rel-to-max %.3e, relative rms %.3e (bf16 activations through up to 99 convs)' % (k, rel, rms)
