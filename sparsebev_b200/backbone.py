"""Backbone conv wrappers on the tcgen05 implicit-GEMM convolution.

Mirrors, with the reference's names / constructor arguments / state-dict keys:
  * `Conv2d`                       /root/reference/models/backbones/eva02/wrappers.py:76-120
                                   (torch.nn.Conv2d + `norm=` + `activation=` keyword arguments, norm before activation);
  * `conv3x3`, `conv1x1`           /root/reference/models/backbones/vovnet.py:117-154
                                   ((name, module) lists of Conv2d + BatchNorm2d + ReLU) and `ConvBNReLUSequence` that runs such a list;
  * `ResNet`, `FPN`                the mmdet 2.28.2 modules the reference builds from configs/r50_nuimg_704x256.py:31-45
                                   (third party, absent here: restated from their documented structure; parity of the
                                   *architecture glue* is therefore unpinned, the convolution arithmetic is pinned against
                                   torch.nn.functional.conv2d in tests/test_gpu_backbone.py);
  * `extract_img_feat`             /root/reference/models/sparsebev.py:46-59,124-131 (backbone -> neck -> [B, T*N, C, H, W]).

Everything between the image and the FPN outputs stays NHWC bf16 on the device; the FPN levels leave as fp32 NHWC, which
is the gather's zero-copy 'nhwc' layout (SparseBEVTransformerDecoder.prepare_feats), so no regroup / permute copy exists
anywhere between the backbone and the decoder.  Inference only (BatchNorm folded from running statistics).  There is no
cuDNN / eager fallback: unsupported configurations raise.
"""
from collections import OrderedDict

import torch
import torch.nn as nn

from . import ops


def fold_bn(bn, conv_bias=None):
    """BatchNorm2d (eval) -> per-channel (scale, shift) fp32: y = scale * conv + shift."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    shift = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    if conv_bias is not None:
        shift = shift + conv_bias.detach().float() * scale
    return scale.contiguous(), shift.contiguous()


def weight_khwc(weight):
    """nn.Conv2d weight [Cout,Cin,KH,KW] -> [Cout,KH,KW,Cin] (the K-major implicit-GEMM operand), fp32 contiguous."""
    return weight.detach().float().permute(0, 2, 3, 1).contiguous()


class _FusedConv:
    """Device-side cache of one conv (+ BN) in the kernel's layout; rebuilt when a parameter changes.

    in_index (optional, list of length Cin_padded): the kernel's input tensor carries Cin_padded >= conv.in_channels channels;
    entry j names the conv input channel that padded channel j feeds (-1: a zero-padding channel, gets zero weights).
    cout_pad (optional): emit cout_pad >= conv.out_channels channels, the extra ones identically zero (zero weights, zero shift).
    Both exist for VoVNet's 160- / 224-channel OSA layers: the tcgen05 conv kernel wants Cin % 64 == 0, so those tensors are
    carried as 192 / 256 channels with zeros in the tail."""

    def __init__(self, conv, bn=None, in_index=None, cout_pad=None):
        self.conv, self.bn, self._key, self._val = conv, bn, None, None
        self.in_index, self.cout_pad = in_index, cout_pad

    def get(self):
        ps = [self.conv.weight, self.conv.bias] + ([self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var] if self.bn is not None else [])
        key = tuple((p.data_ptr(), p._version, str(p.device)) for p in ps if p is not None)
        if key != self._key:
            conv, bn = self.conv, self.bn
            if bn is not None:
                if bn.training:
                    raise RuntimeError('sparsebev_b200 backbone: BatchNorm must be in eval mode (inference only)')
                scale, shift = fold_bn(bn, conv.bias)
            else:
                scale = None
                shift = (conv.bias.detach().float() if conv.bias is not None else torch.zeros(conv.out_channels, device=conv.weight.device)).contiguous()
            wk = weight_khwc(conv.weight)                                       # [Cout, KH, KW, Cin]
            if self.in_index is not None:
                idx = torch.as_tensor(self.in_index, device=wk.device, dtype=torch.long)
                wp = torch.zeros(wk.shape[0], wk.shape[1], wk.shape[2], idx.numel(), device=wk.device, dtype=wk.dtype)
                wp[..., idx >= 0] = wk[..., idx[idx >= 0]]
                wk = wp
            if self.cout_pad is not None and self.cout_pad > wk.shape[0]:
                extra = self.cout_pad - wk.shape[0]
                wk = torch.cat([wk, torch.zeros(extra, *wk.shape[1:], device=wk.device, dtype=wk.dtype)], 0)
                shift = torch.cat([shift, torch.zeros(extra, device=shift.device)]).contiguous()
                if scale is not None:
                    scale = torch.cat([scale, torch.ones(extra, device=scale.device)]).contiguous()
            w = ops.cast_bf16(wk.contiguous())
            self._key, self._val = key, (w, scale, shift)
        return self._val

    def __call__(self, x, relu=False, residual=None, out_f32=False):
        conv = self.conv
        _check_conv(conv, cin=len(self.in_index) if self.in_index is not None else None, cout=self.cout_pad)
        w, scale, shift = self.get()
        return ops.conv2d_nhwc(x, w, shift, scale, stride=conv.stride[0], pad=conv.padding[0], residual=residual, relu=relu, out_f32=out_f32)


def _check_conv(conv, cin=None, cout=None):
    cin = conv.in_channels if cin is None else cin
    cout = conv.out_channels if cout is None else cout
    if conv.groups != 1 or conv.dilation != (1, 1) or conv.stride[0] != conv.stride[1] or conv.padding[0] != conv.padding[1] \
            or conv.kernel_size[0] != conv.kernel_size[1] or conv.stride[0] not in (1, 2) or conv.padding_mode != 'zeros':
        raise NotImplementedError('sparsebev_b200 conv kernel: groups=1, dilation=1, square kernel, stride 1|2, zero padding only '
                                  '(got %r); there is no cuDNN fallback' % (conv,))
    if cin % 64 or cout % 32:
        raise NotImplementedError('sparsebev_b200 conv kernel: in_channels %% 64 == 0 and out_channels %% 32 == 0 required (got %d -> %d)'
                                  % (cin, cout))


def to_nhwc_bf16(x):
    """NCHW float tensor -> NHWC bf16 (boundary conversion for NCHW callers)."""
    return ops.cast_bf16(x.float().permute(0, 2, 3, 1).contiguous())


def to_nchw_f32(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


class Conv2d(torch.nn.Conv2d):
    """Same interface as the reference wrapper (eva02/wrappers.py:76-120): extra keyword arguments `norm` (a
    normalization layer, applied before the activation) and `activation`.  forward(x: NCHW) -> NCHW runs conv + norm +
    activation as ONE tcgen05 kernel when norm is None / BatchNorm2d (eval) and activation is None / ReLU."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop('norm', None)
        activation = kwargs.pop('activation', None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation
        self._fused = None

    def _relu(self):
        act = self.activation
        if act is None:
            return False
        if isinstance(act, nn.ReLU) or act in (torch.relu, torch.nn.functional.relu):
            return True
        raise NotImplementedError('sparsebev_b200 Conv2d: activation must be None or ReLU (got %r)' % (act,))

    def forward_nhwc(self, x, residual=None, out_f32=False):
        if self.norm is not None and not isinstance(self.norm, nn.BatchNorm2d):
            raise NotImplementedError('sparsebev_b200 Conv2d: norm must be None or BatchNorm2d (got %r)' % (self.norm,))
        if self._fused is None or self._fused.bn is not self.norm:
            self._fused = _FusedConv(self, self.norm)
        return self._fused(x, relu=self._relu(), residual=residual, out_f32=out_f32)

    def forward(self, x):
        return to_nchw_f32(self.forward_nhwc(to_nhwc_bf16(x), out_f32=True))


def conv3x3(in_channels, out_channels, module_name, postfix, stride=1, groups=1, kernel_size=3, padding=1):
    """3x3 convolution with padding: [(name, module)] exactly as the reference builds it (vovnet.py:117-135)."""
    return [
        (f'{module_name}_{postfix}/conv',
         nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, groups=groups, bias=False)),
        (f'{module_name}_{postfix}/norm', nn.BatchNorm2d(out_channels)),
        (f'{module_name}_{postfix}/relu', nn.ReLU(inplace=True)),
    ]


def conv1x1(in_channels, out_channels, module_name, postfix, stride=1, groups=1, kernel_size=1, padding=0):
    """1x1 convolution: [(name, module)] exactly as the reference builds it (vovnet.py:138-154)."""
    return [
        (f'{module_name}_{postfix}/conv',
         nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding, groups=groups, bias=False)),
        (f'{module_name}_{postfix}/norm', nn.BatchNorm2d(out_channels)),
        (f'{module_name}_{postfix}/relu', nn.ReLU(inplace=True)),
    ]


class ConvBNReLUSequence(nn.Sequential):
    """nn.Sequential(OrderedDict(conv3x3(...) + conv1x1(...) + ...)) with the reference's module names; forward_nhwc fuses
    every (Conv2d, BatchNorm2d, ReLU) triple into one kernel launch."""

    def __init__(self, named_modules):
        super().__init__(OrderedDict(named_modules))
        mods = list(self.children())
        if len(mods) % 3:
            raise ValueError('expected (conv, norm, relu) triples')
        self._fused = []
        for i in range(0, len(mods), 3):
            c, n, r = mods[i:i + 3]
            if not (isinstance(c, nn.Conv2d) and isinstance(n, nn.BatchNorm2d) and isinstance(r, nn.ReLU)):
                raise ValueError('expected (Conv2d, BatchNorm2d, ReLU) triples')
            self._fused.append(_FusedConv(c, n))

    def forward_nhwc(self, x):
        for f in self._fused:
            x = f(x, relu=True)
        return x

    def forward(self, x):
        return to_nchw_f32(self.forward_nhwc(to_nhwc_bf16(x)))


class Bottleneck(nn.Module):
    """mmdet / torchvision ResNet bottleneck, style='pytorch' (the stride sits on the 3x3 conv); keys conv1..3, bn1..3, downsample.{0,1}."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self._f1, self._f2, self._f3 = _FusedConv(self.conv1, self.bn1), _FusedConv(self.conv2, self.bn2), _FusedConv(self.conv3, self.bn3)
        self._fd = _FusedConv(downsample[0], downsample[1]) if downsample is not None else None

    def forward_nhwc(self, x):
        identity = x if self._fd is None else self._fd(x)
        out = self._f2(self._f1(x, relu=True), relu=True)
        return self._f3(out, relu=True, residual=identity)          # relu(bn3(conv3) + identity) in the conv epilogue


def _reject_unsupported(who, kwargs, training_only=(), checked=None):
    """mmdet config kwargs the inference mirrors do not implement: the ones that only matter for TRAINING (or for how
    weights get initialised -- the caller loads a checkpoint) are accepted and dropped; anything that would change the
    network the config describes raises instead of silently building a different model."""
    import warnings
    checked = checked or {}
    for k, v in kwargs.items():
        if k in training_only:
            if k in ('init_cfg', 'pretrained') and v is not None:
                warnings.warn('%s: %s=%r is not applied by the inference mirror -- load the weights with load_state_dict' % (who, k, v))
            continue
        if k in checked:
            if not checked[k](v):
                raise NotImplementedError('%s: %s=%r is not implemented by the sparsebev_b200 mirror' % (who, k, v))
            continue
        raise NotImplementedError('%s: unsupported argument %s=%r (it would change the architecture or is unknown to the mirror)' % (who, k, v))


class ResNet(nn.Module):
    """ResNet-50 / -101 with mmdet's state-dict keys (conv1, bn1, layer{1..4}.{i}.*) -- the reference's img_backbone
    (configs/r50_nuimg_704x256.py:31-40: depth=50, num_stages=4, out_indices=(0,1,2,3), style='pytorch', norm_eval)."""
    arch_settings = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3)}

    def __init__(self, depth=50, num_stages=4, out_indices=(0, 1, 2, 3), style='pytorch', **kwargs):
        super().__init__()
        if depth not in self.arch_settings or style != 'pytorch':
            raise NotImplementedError('ResNet depth 50 / 101, style pytorch')
        _reject_unsupported('ResNet', kwargs, training_only=('frozen_stages', 'norm_eval', 'with_cp', 'zero_init_residual', 'init_cfg', 'pretrained'),
                            checked={'norm_cfg': lambda v: v is None or str(v.get('type', 'BN')).startswith('BN'),
                                     'dcn': lambda v: v is None, 'stage_with_dcn': lambda v: not any(v), 'plugins': lambda v: v is None,
                                     'deep_stem': lambda v: not v, 'avg_down': lambda v: not v, 'in_channels': lambda v: v == 3,
                                     'base_channels': lambda v: v == 64, 'strides': lambda v: tuple(v) == (1, 2, 2, 2),
                                     'dilations': lambda v: all(d == 1 for d in v), 'conv_cfg': lambda v: v is None,
                                     'stem_channels': lambda v: v in (None, 64)})
        self.out_indices = tuple(out_indices)
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        inplanes = 64
        self.res_layers = []
        for i, blocks in enumerate(self.arch_settings[depth][:num_stages]):
            planes, stride = 64 * 2 ** i, 1 if i == 0 else 2
            layers = []
            for b in range(blocks):
                down = None
                if b == 0 and (stride != 1 or inplanes != planes * 4):
                    down = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False), nn.BatchNorm2d(planes * 4))
                layers.append(Bottleneck(inplanes, planes, stride if b == 0 else 1, down))
                inplanes = planes * 4
            name = 'layer%d' % (i + 1)
            self.add_module(name, nn.Sequential(*layers))
            self.res_layers.append(name)
        self._stem_key, self._stem_val = None, None

    def _stem(self):
        ps = [self.conv1.weight, self.bn1.weight, self.bn1.bias, self.bn1.running_mean, self.bn1.running_var]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._stem_key:
            scale, shift = fold_bn(self.bn1)
            w = self.conv1.weight.detach().float().permute(2, 3, 1, 0).contiguous()        # [7,7,3,64]
            self._stem_key, self._stem_val = key, (w, scale, shift)
        return self._stem_val

    @torch.no_grad()
    def forward_nhwc(self, img):
        """img NCHW fp32 [N,3,H,W] -> tuple of NHWC bf16 stage outputs (strides 4, 8, 16, 32)."""
        if self.training:
            raise RuntimeError('sparsebev_b200 ResNet: inference only (call .eval())')
        w, scale, shift = self._stem()
        x = ops.maxpool3x3s2_nhwc(ops.stem_conv(img.float().contiguous(), w, scale, shift))
        outs = []
        for i, name in enumerate(self.res_layers):
            for block in getattr(self, name):
                x = block.forward_nhwc(x)
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)

    def forward(self, img):
        return tuple(to_nchw_f32(o) for o in self.forward_nhwc(img))


class _ConvModule(nn.Module):
    """mmcv ConvModule without norm / activation: keeps the `.conv.weight` / `.conv.bias` key names."""

    def __init__(self, cin, cout, k, padding=0):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=padding)
        self._fused = _FusedConv(self.conv)


class FPN(nn.Module):
    """mmdet FPN (configs/r50_nuimg_704x256.py:41-45: in_channels=[256,512,1024,2048], out_channels=256, num_outs=4; the r101
    config asks for num_outs=5): lateral 1x1 convs, nearest top-down path, 3x3 output convs, extra levels by stride-2
    subsampling (F.max_pool2d(x, 1, stride=2)).  The top-down add is the residual operand of the lateral conv's epilogue."""

    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False, **kwargs):
        super().__init__()
        if add_extra_convs or end_level != -1:
            raise NotImplementedError('FPN: add_extra_convs / end_level are not used by the reference configs')
        _reject_unsupported('FPN', kwargs, training_only=('init_cfg',),
                            checked={'relu_before_extra_convs': lambda v: not v, 'no_norm_on_lateral': lambda v: True, 'conv_cfg': lambda v: v is None,
                                     'norm_cfg': lambda v: v is None, 'act_cfg': lambda v: v is None,
                                     'upsample_cfg': lambda v: v is None or (v.get('mode', 'nearest') == 'nearest' and 'scale_factor' not in v)})
        self.in_channels, self.out_channels, self.num_outs, self.start_level = list(in_channels), out_channels, num_outs, start_level
        self.lateral_convs = nn.ModuleList(_ConvModule(c, out_channels, 1) for c in self.in_channels[start_level:])
        self.fpn_convs = nn.ModuleList(_ConvModule(out_channels, out_channels, 3, padding=1) for _ in self.in_channels[start_level:])

    @torch.no_grad()
    def forward_nhwc(self, inputs):
        """inputs: NHWC bf16 stage outputs -> list of num_outs NHWC fp32 levels."""
        feats = list(inputs)[self.start_level:]
        n = len(feats)
        lats = [None] * n
        for i in range(n - 1, -1, -1):                      # top-down: lateral(i) + nearest-upsampled lateral(i+1), one kernel
            lats[i] = self.lateral_convs[i]._fused(feats[i], residual=lats[i + 1] if i + 1 < n else None)
        outs = [self.fpn_convs[i]._fused(lats[i], out_f32=True) for i in range(n)]
        while len(outs) < self.num_outs:
            outs.append(ops.subsample2_nhwc(outs[-1]))
        return outs

    def forward(self, inputs):
        return tuple(o.permute(0, 3, 1, 2) for o in self.forward_nhwc([to_nhwc_bf16(x) for x in inputs]))


@torch.no_grad()
def extract_img_feat(backbone, neck, img):
    """img [B, T*N, 3, H, W] -> list of [B, T*N, C, H', W'] fp32 whose memory is channels-last, i.e. exactly what
    SparseBEVTransformerDecoder.prepare_feats consumes without a copy (reference: models/sparsebev.py:46-59,124-131)."""
    B, TN = img.shape[:2]
    stages = backbone.forward_nhwc(img.reshape(B * TN, *img.shape[2:]))
    if isinstance(stages, dict):                    # VoVNet returns a dict (models/sparsebev.py:52-53: list(x.values()))
        stages = list(stages.values())
    levels = neck.forward_nhwc(stages)
    return [f.view(B, TN, *f.shape[1:]).permute(0, 1, 4, 2, 3) for f in levels]


# ------------------------------------------------------------------------------------------------
# VoVNet (V-99-eSE and the other non-depthwise specs): /root/reference/models/backbones/vovnet.py:12-90 (specs), :157-178 (eSE),
# :181-239 (OSA module / stage), :243-359 (network).  Same module tree and state-dict keys (`stem.stem_1/conv.weight`,
# `stage3.OSA3_2.layers.4.OSA3_2_4/norm.running_var`, `stage5.OSA5_3.ese.fc.bias`, ...); inference only.
_VOV_SPECS = {
    'V-19-slim-eSE': dict(stem=[64, 64, 128], stage_conv_ch=[64, 80, 96, 112], stage_out_ch=[112, 256, 384, 512], layer_per_block=3, block_per_stage=[1, 1, 1, 1]),
    'V-19-eSE': dict(stem=[64, 64, 128], stage_conv_ch=[128, 160, 192, 224], stage_out_ch=[256, 512, 768, 1024], layer_per_block=3, block_per_stage=[1, 1, 1, 1]),
    'V-39-eSE': dict(stem=[64, 64, 128], stage_conv_ch=[128, 160, 192, 224], stage_out_ch=[256, 512, 768, 1024], layer_per_block=5, block_per_stage=[1, 1, 2, 2]),
    'V-57-eSE': dict(stem=[64, 64, 128], stage_conv_ch=[128, 160, 192, 224], stage_out_ch=[256, 512, 768, 1024], layer_per_block=5, block_per_stage=[1, 1, 4, 3]),
    'V-99-eSE': dict(stem=[64, 64, 128], stage_conv_ch=[128, 160, 192, 224], stage_out_ch=[256, 512, 768, 1024], layer_per_block=5, block_per_stage=[1, 3, 9, 3]),
}


def _pad64(c):
    return (c + 63) // 64 * 64


class eSEModule(nn.Module):
    """vovnet.py:166-178 (parameters: fc = 1x1 conv with bias)."""

    def __init__(self, channel, reduction=4):
        super().__init__()
        self.fc = nn.Conv2d(channel, channel, kernel_size=1, padding=0)

    def forward_nhwc(self, x, identity=None):
        C = self.fc.out_channels
        return ops.ese_nhwc(x, self.fc.weight.detach().float().reshape(C, C).contiguous(), self.fc.bias.detach().float().contiguous(), identity)


class _OSA_module(nn.Module):
    """One-shot-aggregation block (vovnet.py:181-225): layer_per_block 3x3 convs in sequence, all intermediate outputs (and the
    input) concatenated, 1x1 conv, eSE, optional identity.  Channel counts that are not multiples of 64 (160, 224) travel
    zero-padded to the next multiple; the consumers' weights are laid out for the padded tensors (_FusedConv.in_index)."""

    def __init__(self, in_ch, stage_ch, concat_ch, layer_per_block, module_name, SE=False, identity=False, depthwise=False, with_cp=False):
        super().__init__()
        if depthwise:
            raise NotImplementedError('sparsebev_b200 VoVNet: depthwise (dw) specs are not implemented')
        self.identity = identity
        self.layers = nn.ModuleList()
        sp = _pad64(stage_ch)
        self._fused = []
        cin, cin_carried = in_ch, _pad64(in_ch)
        for i in range(layer_per_block):
            seq = nn.Sequential(OrderedDict(conv3x3(cin, stage_ch, module_name, i)))
            self.layers.append(seq)
            index = list(range(cin)) + [-1] * (cin_carried - cin)
            self._fused.append(_FusedConv(seq[0], seq[1], in_index=index if cin_carried != cin else None, cout_pad=sp if sp != stage_ch else None))
            cin, cin_carried = stage_ch, sp
        self.concat = nn.Sequential(OrderedDict(conv1x1(in_ch + layer_per_block * stage_ch, concat_ch, module_name, 'concat')))
        index = list(range(in_ch)) + [-1] * (_pad64(in_ch) - in_ch)
        for i in range(layer_per_block):
            index += list(range(in_ch + i * stage_ch, in_ch + (i + 1) * stage_ch)) + [-1] * (sp - stage_ch)
        plain = index == list(range(len(index)))
        self._fused_concat = _FusedConv(self.concat[0], self.concat[1], in_index=None if plain else index)
        self.ese = eSEModule(concat_ch)

    def forward_nhwc(self, x):
        outs, h = [x], x
        for f in self._fused:
            h = f(h, relu=True)
            outs.append(h)
        xt = self._fused_concat(torch.cat(outs, dim=-1), relu=True)            # (the concatenation is a plain NHWC copy)
        return self.ese.forward_nhwc(xt, x if self.identity else None)


class _OSA_stage(nn.Sequential):
    """vovnet.py:228-262: MaxPool2d(3, 2, ceil_mode=True) (all stages but stage2) + block_per_stage OSA modules."""

    def __init__(self, in_ch, stage_ch, concat_ch, block_per_stage, layer_per_block, stage_num, SE=False, depthwise=False, with_cp=False):
        super().__init__()
        if stage_num != 2:
            self.add_module('Pooling', nn.MaxPool2d(kernel_size=3, stride=2, ceil_mode=True))
        name = 'OSA%d_1' % stage_num
        self.add_module(name, _OSA_module(in_ch, stage_ch, concat_ch, layer_per_block, name, SE, depthwise=depthwise))
        for i in range(block_per_stage - 1):
            name = 'OSA%d_%d' % (stage_num, i + 2)
            self.add_module(name, _OSA_module(concat_ch, stage_ch, concat_ch, layer_per_block, name, SE, identity=True, depthwise=depthwise))

    def forward_nhwc(self, x):
        for m in self.children():
            x = ops.maxpool3x3s2_ex_nhwc(x, pad=0, ceil_mode=True) if isinstance(m, nn.MaxPool2d) else m.forward_nhwc(x)
        return x


class VoVNet(nn.Module):
    """The reference's img_backbone for BASELINE config 5 (configs/vov99_dd3d_1600x640_trainval_future.py:14-22:
    spec_name='V-99-eSE', out_features=['stage2', ..., 'stage5']), same constructor arguments and state-dict keys as
    /root/reference/models/backbones/vovnet.py:265-359."""

    def __init__(self, spec_name, input_ch=3, out_features=None, frozen_stages=-1, norm_eval=True, with_cp=False, pretrained=None, init_cfg=None):
        super().__init__()
        if spec_name not in _VOV_SPECS:
            raise NotImplementedError('sparsebev_b200 VoVNet: spec %r (the depthwise specs are not implemented)' % (spec_name,))
        if input_ch != 3:
            raise NotImplementedError('sparsebev_b200 VoVNet: 3 input channels')
        spec = _VOV_SPECS[spec_name]
        stem_ch, stage_ch, concat_ch = spec['stem'], spec['stage_conv_ch'], spec['stage_out_ch']
        self._out_features = list(out_features) if out_features is not None else ['stage2', 'stage3', 'stage4', 'stage5']
        stem = conv3x3(input_ch, stem_ch[0], 'stem', '1', 2) + conv3x3(stem_ch[0], stem_ch[1], 'stem', '2', 1) + conv3x3(stem_ch[1], stem_ch[2], 'stem', '3', 2)
        self.add_module('stem', nn.Sequential(OrderedDict(stem)))
        if stem_ch[0] != 64:
            raise NotImplementedError('sparsebev_b200 VoVNet: the 3-channel stem kernel emits 64 channels')
        self._stem2, self._stem3 = _FusedConv(self.stem[3], self.stem[4]), _FusedConv(self.stem[6], self.stem[7])
        in_ch = [stem_ch[2]] + concat_ch[:-1]
        self.stage_names = []
        for i in range(4):
            name = 'stage%d' % (i + 2)
            self.stage_names.append(name)
            self.add_module(name, _OSA_stage(in_ch[i], stage_ch[i], concat_ch[i], spec['block_per_stage'][i], spec['layer_per_block'], i + 2, True, False))
        self._stem_key, self._stem_val = None, None

    def _stem1(self):
        conv, bn = self.stem[0], self.stem[1]
        ps = [conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        key = tuple((p.data_ptr(), p._version) for p in ps)
        if key != self._stem_key:
            scale, shift = fold_bn(bn)
            self._stem_key, self._stem_val = key, (conv.weight.detach().float().permute(2, 3, 1, 0).contiguous(), scale, shift)   # [3,3,3,64]
        return self._stem_val

    @torch.no_grad()
    def forward_nhwc(self, img):
        """img NCHW fp32 [N,3,H,W] -> dict name -> NHWC bf16 (the reference returns a dict too, vovnet.py:327-337)."""
        if self.training:
            raise RuntimeError('sparsebev_b200 VoVNet: inference only (call .eval())')
        w, scale, shift = self._stem1()
        x = ops.stem_conv(img.float().contiguous(), w, scale, shift)
        x = self._stem3(self._stem2(x, relu=True), relu=True)
        outs = {}
        if 'stem' in self._out_features:
            outs['stem'] = x
        for name in self.stage_names:
            x = getattr(self, name).forward_nhwc(x)
            if name in self._out_features:
                outs[name] = x
        return outs

    def forward(self, img):
        return {k: to_nchw_f32(v) for k, v in self.forward_nhwc(img).items()}


def enable(force=False):
    """Register the inference-only mirrors with mmdet's registries (when mmdet is importable): always under the distinct names
    `ResNetB200` / `FPNB200`; with force=True ALSO under mmdet's own names `ResNet` / `FPN`, replacing mmdet's classes for
    every model built afterwards in this process -- an explicit opt-in, because the mirrors cannot train and reject the
    architecture-changing kwargs they do not implement (see _reject_unsupported)."""
    from mmdet.models.builder import BACKBONES, NECKS
    BACKBONES.register_module(name='ResNetB200', module=ResNet, force=True)
    BACKBONES.register_module(name='VoVNetB200', module=VoVNet, force=True)
    NECKS.register_module(name='FPNB200', module=FPN, force=True)
    if force:
        BACKBONES.register_module(module=VoVNet, force=True)
        BACKBONES.register_module(module=ResNet, force=True)
        NECKS.register_module(module=FPN, force=True)


try:                                    # distinct names only: importing this module never shadows mmdet's own ResNet / FPN
    enable(force=False)
except Exception:                       # pragma: no cover  (mmdet absent in the build image)
    pass
