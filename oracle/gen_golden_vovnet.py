"""Generate tests/golden/vovnet.npz by running the REAL reference VoVNet -- TEST INFRASTRUCTURE.

/root/reference/models/backbones/vovnet.py is executed UNMODIFIED (imported in place) on the CPU in fp32; its two third-party
imports are stubbed (mmcv.runner.BaseModule -> nn.Module storing init_cfg, mmdet.models.builder.BACKBONES -> a registry whose
register_module() is the identity decorator).  The 69.5 M parameters of V-99-eSE are not stored: `seeded_init` fills any
module with the same state-dict layout reproducibly, so the tests re-create them.  The fixture holds the reference's
state-dict key list (the mirror must have exactly these keys and shapes) and its four stage outputs for a small odd-sized
input (exercises MaxPool2d(ceil_mode=True) overhang).

Run in the build container only: `python oracle/gen_golden_vovnet.py`.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'tests', 'golden')
REF = '/root/reference/models/backbones/vovnet.py'


def seeded_init(module, seed=0):
    """Deterministic 'trained-like' parameters for any module, independent of construction order: every tensor of the
    state dict is drawn from its own generator seeded by (seed, position): conv / fc weights ~ N(0, sqrt(2 / fan_in)) (so that
    activations keep unit scale through ~100 layers), BN weight ~ U(0.8, 1.2), BN bias / running_mean ~ N(0, 0.1), running_var ~
    U(0.8, 1.2), biases ~ N(0, 0.1)."""
    sd = module.state_dict()
    for i, (k, v) in enumerate(sd.items()):
        g = torch.Generator().manual_seed(seed * 100003 + i)
        if k.endswith('num_batches_tracked'):
            continue
        if v.dim() == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            v.copy_(torch.randn(v.shape, generator=g) * (2.0 / fan_in) ** 0.5)
        elif k.endswith('running_var') or (k.endswith('weight') and v.dim() == 1):
            v.copy_(0.8 + 0.4 * torch.rand(v.shape, generator=g))
        else:
            v.copy_(0.1 * torch.randn(v.shape, generator=g))
    module.load_state_dict(sd)
    return module


def import_reference_vovnet():
    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        m.__dict__.update(attrs)
        sys.modules[name] = m
    mod('mmcv'); mod('mmcv.runner', BaseModule=BaseModule)
    mod('mmdet'); mod('mmdet.models'); mod('mmdet.models.builder', BACKBONES=_Registry())
    spec = importlib.util.spec_from_file_location('ref_vovnet', REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_image(seed=5, n=2, h=100, w=164):
    return torch.randn(n, 3, h, w, generator=torch.Generator().manual_seed(seed))


def main():
    torch.set_num_threads(8)
    ref = import_reference_vovnet()
    feats = ['stage2', 'stage3', 'stage4', 'stage5']
    net = ref.VoVNet('V-99-eSE', out_features=feats, norm_eval=True, frozen_stages=1, with_cp=True)
    seeded_init(net, seed=3).eval()
    img = test_image()
    with torch.no_grad():
        out = net(img)
    keys = list(net.state_dict().keys())
    shapes = [list(v.shape) for v in net.state_dict().values()]
    np.savez_compressed(os.path.join(OUT, 'vovnet.npz'), keys=np.array(keys), shapes=np.array([str(s) for s in shapes]),
                        **{k: out[k].numpy() for k in feats})
    for k in feats:
        print(k, tuple(out[k].shape), 'rms %.3f' % float(out[k].pow(2).mean().sqrt()))
    print('wrote', os.path.join(OUT, 'vovnet.npz'), len(keys), 'keys')


if __name__ == '__main__':
    main()
