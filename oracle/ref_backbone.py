"""TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's checker legs, never by sparsebev_b200/).

Plain-PyTorch fp32 restatement of the reference's image branch:
  * conv + norm + activation as the reference's wrappers compute it -- F.conv2d -> BatchNorm2d (eval) -> ReLU
    (/root/reference/models/backbones/eva02/wrappers.py:111-119, models/backbones/vovnet.py:117-154);
  * the mmdet 2.28.2 ResNet (style='pytorch') and FPN forward passes the reference builds from
    configs/r50_nuimg_704x256.py:31-45.  mmdet is absent from this image (no network), so these two are restated from
    its documented structure: **parity unpinned** for the architecture glue; the convolution arithmetic itself is
    torch.nn.functional.conv2d, i.e. the very call the reference makes.
`emulate_bf16=True` rounds weights and every conv input to bf16 (what the tcgen05 path feeds its tensor cores; the
reference runs this branch under fp16 autocast, models/sparsebev.py:46) so the comparison isolates kernel errors from
the operand rounding both share.
"""
import torch
import torch.nn.functional as F


def _r(x, on):
    return x.bfloat16().float() if on else x


def conv_bn_act(x, weight, bias=None, bn=None, relu=False, stride=1, padding=0, residual=None, emulate_bf16=False):
    """x NCHW fp32.  bn = (gamma, beta, running_mean, running_var, eps) or None.  residual NCHW (added before the ReLU;
    nearest-upsampled to the output size when smaller, as mmdet FPN's top-down path does)."""
    y = F.conv2d(_r(x, emulate_bf16), _r(weight, emulate_bf16), None, stride, padding)
    if bn is not None:
        g, b, m, v, eps = bn
        scale = g / torch.sqrt(v + eps)
        shift = b - m * scale
        if bias is not None:
            shift = shift + bias * scale
        y = y * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    elif bias is not None:
        y = y + bias.view(1, -1, 1, 1)
    if residual is not None:
        r = _r(residual, emulate_bf16)
        if r.shape[-2:] != y.shape[-2:]:
            r = F.interpolate(r, size=y.shape[-2:], mode='nearest')
        y = y + r
    return torch.relu(y) if relu else y


def _bn(sd, prefix):
    return (sd[prefix + '.weight'], sd[prefix + '.bias'], sd[prefix + '.running_mean'], sd[prefix + '.running_var'], 1e-5)


def resnet_forward(img, sd, depth=50, emulate_bf16=False):
    """img NCHW fp32, sd = state dict with mmdet / torchvision ResNet keys -> [C2, C3, C4, C5] NCHW fp32."""
    blocks = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}[depth]
    e = emulate_bf16
    x = conv_bn_act(img, sd['conv1.weight'], bn=_bn(sd, 'bn1'), relu=True, stride=2, padding=3)      # stem stays fp32
    x = _r(x, e)
    x = F.max_pool2d(x, 3, 2, 1)
    outs = []
    for li, nb in enumerate(blocks):
        for b in range(nb):
            p = 'layer%d.%d.' % (li + 1, b)
            stride = 2 if (b == 0 and li > 0) else 1
            identity = x
            if p + 'downsample.0.weight' in sd:
                identity = _r(conv_bn_act(x, sd[p + 'downsample.0.weight'], bn=_bn(sd, p + 'downsample.1'), stride=stride, emulate_bf16=e), e)
            o = _r(conv_bn_act(x, sd[p + 'conv1.weight'], bn=_bn(sd, p + 'bn1'), relu=True, emulate_bf16=e), e)
            o = _r(conv_bn_act(o, sd[p + 'conv2.weight'], bn=_bn(sd, p + 'bn2'), relu=True, stride=stride, padding=1, emulate_bf16=e), e)
            x = _r(conv_bn_act(o, sd[p + 'conv3.weight'], bn=_bn(sd, p + 'bn3'), relu=True, residual=identity, emulate_bf16=e), e)
        outs.append(x)
    return outs


def fpn_forward(feats, sd, num_outs, emulate_bf16=False):
    """mmdet FPN: laterals, nearest top-down adds, 3x3 output convs, extra levels by max_pool2d(k=1, s=2)."""
    e = emulate_bf16
    n = len(feats)
    lats = [None] * n
    for i in range(n - 1, -1, -1):
        lats[i] = _r(conv_bn_act(feats[i], sd['lateral_convs.%d.conv.weight' % i], sd['lateral_convs.%d.conv.bias' % i],
                                 residual=lats[i + 1] if i + 1 < n else None, emulate_bf16=e), e)
    outs = [conv_bn_act(lats[i], sd['fpn_convs.%d.conv.weight' % i], sd['fpn_convs.%d.conv.bias' % i], padding=1, emulate_bf16=e) for i in range(n)]
    while len(outs) < num_outs:
        outs.append(F.max_pool2d(outs[-1], 1, stride=2))
    return outs
