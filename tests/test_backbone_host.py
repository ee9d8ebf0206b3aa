"""Host-side logic of the backbone conv path (no GPU): BatchNorm folding, weight relayout, module / state-dict key names
of the ResNet / FPN / vovnet-style mirrors, and the oracle's own consistency with torch modules."""
import torch

from oracle import ref_backbone as RB
from sparsebev_b200 import backbone as BB


def test_fold_bn_matches_torch_batchnorm_eval():
    g = torch.Generator().manual_seed(0)
    bn = torch.nn.BatchNorm2d(16).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(16, generator=g) + 0.5); bn.bias.copy_(torch.randn(16, generator=g))
        bn.running_mean.copy_(torch.randn(16, generator=g)); bn.running_var.copy_(torch.rand(16, generator=g) + 0.5)
    x = torch.randn(2, 16, 5, 7, generator=g)
    bias = torch.randn(16, generator=g)
    scale, shift = BB.fold_bn(bn, bias)
    with torch.no_grad():
        want = bn(x + bias.view(1, -1, 1, 1))
    assert torch.allclose(x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1), want, rtol=1e-5, atol=1e-5)


def test_weight_relayout_is_tap_major_channel_minor():
    w = torch.arange(2 * 3 * 2 * 2, dtype=torch.float32).reshape(2, 3, 2, 2)
    k = BB.weight_khwc(w)
    assert k.shape == (2, 2, 2, 3) and k.is_contiguous()
    assert float(k[1, 0, 1, 2]) == float(w[1, 2, 0, 1])


def test_resnet_and_fpn_state_dict_keys_follow_mmdet():
    net = BB.ResNet(depth=50)
    keys = set(net.state_dict().keys())
    for k in ['conv1.weight', 'bn1.running_var', 'layer1.0.conv1.weight', 'layer1.0.downsample.0.weight', 'layer1.0.downsample.1.running_mean',
              'layer2.3.bn3.weight', 'layer3.5.conv2.weight', 'layer4.2.conv3.weight', 'layer4.0.downsample.1.bias']:
        assert k in keys, k
    assert 'layer1.1.downsample.0.weight' not in keys and 'fc.weight' not in keys
    assert net.layer2[0].conv2.stride == (2, 2) and net.layer2[0].conv1.stride == (1, 1)          # style='pytorch'
    assert sum(p.numel() for p in net.parameters()) == 23508032                                  # torchvision resnet50 minus fc
    assert len(BB.ResNet(depth=101).layer3) == 23
    neck = BB.FPN([256, 512, 1024, 2048], 256, 4)
    nk = set(neck.state_dict().keys())
    assert nk == {'%s.%d.conv.%s' % (a, i, b) for a in ('lateral_convs', 'fpn_convs') for i in range(4) for b in ('weight', 'bias')}
    assert neck.fpn_convs[0].conv.padding == (1, 1) and neck.lateral_convs[3].conv.in_channels == 2048


def test_vovnet_style_builders_keep_reference_names():
    names = [n for n, _ in BB.conv3x3(64, 128, 'OSA2_1', 0)] + [n for n, _ in BB.conv1x1(128, 64, 'OSA2_1', 'concat')]
    assert names == ['OSA2_1_0/conv', 'OSA2_1_0/norm', 'OSA2_1_0/relu', 'OSA2_1_concat/conv', 'OSA2_1_concat/norm', 'OSA2_1_concat/relu']
    seq = BB.ConvBNReLUSequence(BB.conv3x3(64, 128, 'OSA2_1', 0))
    assert list(seq.state_dict().keys())[0] == 'OSA2_1_0/conv.weight'
    c = BB.Conv2d(8, 16, 3, padding=1, norm=torch.nn.BatchNorm2d(16), activation=torch.nn.ReLU())
    assert isinstance(c, torch.nn.Conv2d) and c.norm is not None and c.activation is not None


def test_product_path_has_no_cpu_fallback():
    import pytest
    c = BB.Conv2d(64, 64, 1)
    with pytest.raises((RuntimeError, NotImplementedError)):
        c(torch.randn(1, 64, 4, 4))                                     # CPU tensor: raises, never falls back to F.conv2d


def test_oracle_resnet_fpn_matches_torch_modules():
    """The oracle's functional ResNet / FPN == running the mirrored nn.Modules' parameters through torch layers."""
    torch.manual_seed(0)
    net = BB.ResNet(depth=50).eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.75, 1.25)
    img = torch.randn(1, 3, 64, 64)
    outs = RB.resnet_forward(img, net.state_dict(), 50)
    assert [tuple(o.shape[1:]) for o in outs] == [(256, 16, 16), (512, 8, 8), (1024, 4, 4), (2048, 2, 2)]
    with torch.no_grad():                                              # torch module path for the first stage
        x = net.maxpool(net.relu(net.bn1(net.conv1(img))))
        for blk in net.layer1:
            idt = x if blk.downsample is None else blk.downsample(x)
            o = blk.relu(blk.bn1(blk.conv1(x))); o = blk.relu(blk.bn2(blk.conv2(o)))
            x = blk.relu(blk.bn3(blk.conv3(o)) + idt)
    assert torch.allclose(outs[0], x, rtol=1e-4, atol=1e-5)
    neck = BB.FPN([256, 512, 1024, 2048], 256, 5).eval()
    f = RB.fpn_forward(outs, neck.state_dict(), 5)
    assert [tuple(o.shape[1:]) for o in f] == [(256, 16, 16), (256, 8, 8), (256, 4, 4), (256, 2, 2), (256, 1, 1)]


def test_reference_config_kwargs_accepted_and_architecture_changes_rejected():
    """The reference's own backbone / neck config (configs/r50_nuimg_704x256.py:31-45, r101_nuimg_1408x512.py:14-25) builds;
    training-only knobs are dropped, anything that would silently change the network raises."""
    import pytest
    import warnings
    net = BB.ResNet(depth=101, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1, norm_cfg=dict(type='BN2d', requires_grad=True),
                    norm_eval=True, style='pytorch', with_cp=True)
    assert len(net.layer3) == 23
    BB.FPN(in_channels=[256, 512, 1024, 2048], out_channels=256, num_outs=5)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        BB.ResNet(depth=50, init_cfg=dict(type='Pretrained', checkpoint='x.pth'))
    assert any('load the weights' in str(x.message) for x in w)
    for bad in (dict(dcn=dict(type='DCNv2')), dict(norm_cfg=dict(type='GN', num_groups=32)), dict(deep_stem=True), dict(some_new_option=1)):
        with pytest.raises(NotImplementedError):
            BB.ResNet(depth=50, **bad)
    for bad in (dict(relu_before_extra_convs=True), dict(upsample_cfg=dict(mode='bilinear')), dict(norm_cfg=dict(type='BN'))):
        with pytest.raises(NotImplementedError):
            BB.FPN([256, 512], 256, 2, **bad)


def test_sampling_shims_refuse_to_drop_gradients():
    """sparsebev_b200.sampling.{sampling_4d, make_sample_points} are forward-only: with grad-requiring inputs they raise
    (instead of silently returning a tensor without grad_fn) before touching the GPU."""
    import pytest
    from sparsebev_b200 import sampling
    pts = torch.zeros(1, 2, 1, 4, 4, 3, requires_grad=True)
    with pytest.raises(RuntimeError, match='forward-only'):
        sampling.sampling_4d(pts, [torch.zeros(4, 6, 2, 2, 64)], torch.zeros(1, 2, 4, 1, 4, 1), torch.zeros(1, 6, 4, 4), 8, 8)
    feats = [torch.zeros(4, 6, 2, 2, 64, requires_grad=True)]
    with pytest.raises(RuntimeError, match='forward-only'):
        sampling.sampling_4d(pts.detach(), feats, torch.zeros(1, 2, 4, 1, 4, 1), torch.zeros(1, 6, 4, 4), 8, 8)
    with pytest.raises(RuntimeError, match='forward-only'):
        sampling.make_sample_points(torch.zeros(1, 2, 10), torch.zeros(1, 2, 16, 3, requires_grad=True), [-1, -1, -1, 1, 1, 1])


def test_vovnet_mirror_has_the_reference_state_dict(golden_dir):
    """Key names and shapes of the VoVNet mirror equal those of the REAL reference class (tests/golden/vovnet.npz, written by
    oracle/gen_golden_vovnet.py from /root/reference/models/backbones/vovnet.py): reference checkpoints load with strict=True."""
    import os
    import numpy as np
    g = np.load(os.path.join(golden_dir, 'vovnet.npz'))
    net = BB.VoVNet('V-99-eSE', out_features=['stage2', 'stage3', 'stage4', 'stage5'], norm_eval=True, frozen_stages=1, with_cp=True)
    sd = net.state_dict()
    assert list(sd.keys()) == [str(k) for k in g['keys']]
    assert [str(list(v.shape)) for v in sd.values()] == [str(s) for s in g['shapes']]
    assert sum(p.numel() for p in net.parameters()) == 69523520
    # padded-channel bookkeeping of a 160-channel OSA block: 256 + 5 x 160 concat inputs carried as 256 + 5 x 192
    osa = net.stage3.OSA3_1
    assert osa._fused[1].in_index == list(range(160)) + [-1] * 32 and osa._fused[1].cout_pad == 192
    idx = osa._fused_concat.in_index
    assert len(idx) == 256 + 5 * 192 and [i for i in idx if i >= 0] == list(range(256 + 5 * 160))
    assert net.stage4.OSA4_1._fused_concat.in_index is None                 # 192-channel stages need no padding
