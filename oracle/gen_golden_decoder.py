"""Generate tests/golden/decoder.npz by running the REAL reference decoder -- TEST INFRASTRUCTURE.

/root/reference/models/sparsebev_transformer.py is executed UNMODIFIED (imported in place, never copied) on the CPU:
SparseBEVTransformer -> SparseBEVTransformerDecoder -> 6x SparseBEVTransformerDecoderLayer with the real
position_encoder / SparseBEVSelfAttention (calc_bbox_dists, tau bias) / SparseBEVSampling (motion warp, make_sample_points,
sampling_4d, msmv_sampling_pytorch) / AdaptiveMixing / refine_bbox / velocity rescale / metadata handling / feature regroup.

Its third-party imports are absent here (mmcv-full 1.6.0, mmdet 2.28.2; no network), so exactly FOUR names are stubbed:
  mmcv.runner.BaseModule                          nn.Module that stores init_cfg
  mmcv.cnn.bias_init_with_prob                    -log((1 - p) / p)
  mmcv.cnn.bricks.transformer.MultiheadAttention  restated from mmcv 1.6.0: `.attn = nn.MultiheadAttention(embed_dims, num_heads,
                                                  attn_drop)`, key = value = identity = query, (batch_first) transposes around the
                                                  call, returns identity + dropout_layer(proj_drop(out))
  mmcv.cnn.bricks.transformer.FFN                 restated from mmcv 1.6.0: `.layers = Sequential(Sequential(Linear, ReLU, Dropout),
                                                  Linear, Dropout)`, returns identity + dropout_layer(layers(x))
  mmdet.models.utils.builder.TRANSFORMER          registry whose register_module() is the identity decorator
Those two mmcv wrappers therefore stay "parity unpinned" (DESIGN.md section 2); everything else in the decoder is pinned by
this fixture: tests/test_oracle_golden.py::test_decoder_restatement_matches_real_reference_decoder holds
oracle/ref_torch.py::decoder to these outputs, and the GPU tests hold the CUDA decoder to oracle/ref_torch.py.

Run in the build container only: `python oracle/gen_golden_decoder.py`.
"""
import copy
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.gen_golden import import_reference, OUT, REF      # noqa: E402
from sparsebev_b200 import synthetic as S                     # noqa: E402  (input generators only: no kernels involved)


class BaseModule(nn.Module):
    def __init__(self, init_cfg=None):
        super().__init__()
        self.init_cfg = init_cfg


class MultiheadAttention(BaseModule):
    """mmcv 1.6.0 mmcv/cnn/bricks/transformer.py::MultiheadAttention, the parts the reference exercises."""

    def __init__(self, embed_dims, num_heads, attn_drop=0., proj_drop=0., dropout_layer=None, init_cfg=None, batch_first=False, **kwargs):
        super().__init__(init_cfg)
        self.embed_dims, self.num_heads, self.batch_first = embed_dims, num_heads, batch_first
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, attn_drop, **kwargs)
        self.proj_drop = nn.Dropout(proj_drop)
        self.dropout_layer = nn.Dropout(dropout_layer['drop_prob']) if dropout_layer else nn.Identity()

    def forward(self, query, key=None, value=None, identity=None, query_pos=None, key_pos=None, attn_mask=None, key_padding_mask=None, **kwargs):
        key = query if key is None else key
        value = key if value is None else value
        identity = query if identity is None else identity
        if query_pos is not None:
            query = query + query_pos
        if key_pos is not None:
            key = key + key_pos
        if self.batch_first:
            query, key, value = query.transpose(0, 1), key.transpose(0, 1), value.transpose(0, 1)
        out = self.attn(query=query, key=key, value=value, attn_mask=attn_mask, key_padding_mask=key_padding_mask)[0]
        if self.batch_first:
            out = out.transpose(0, 1)
        return identity + self.dropout_layer(self.proj_drop(out))


class FFN(BaseModule):
    """mmcv 1.6.0 mmcv/cnn/bricks/transformer.py::FFN with its defaults (num_fcs=2, ReLU, add_identity=True)."""

    def __init__(self, embed_dims=256, feedforward_channels=1024, num_fcs=2, act_cfg=None, ffn_drop=0., dropout_layer=None, add_identity=True, init_cfg=None, **kwargs):
        super().__init__(init_cfg)
        layers, in_channels = [], embed_dims
        for _ in range(num_fcs - 1):
            layers.append(nn.Sequential(nn.Linear(in_channels, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)))
            in_channels = feedforward_channels
        layers.append(nn.Linear(feedforward_channels, embed_dims))
        layers.append(nn.Dropout(ffn_drop))
        self.layers = nn.Sequential(*layers)
        self.dropout_layer = nn.Dropout(dropout_layer['drop_prob']) if dropout_layer else nn.Identity()
        self.add_identity = add_identity

    def forward(self, x, identity=None):
        out = self.layers(x)
        if not self.add_identity:
            return self.dropout_layer(out)
        return (x if identity is None else identity) + self.dropout_layer(out)


class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


def import_reference_decoder():
    import_reference()                       # models, models.bbox.utils, models.utils, models.csrc.wrapper, models.sparsebev_sampling

    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__path__ = []
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    mod('mmcv')
    mod('mmcv.runner', BaseModule=BaseModule)
    mod('mmcv.cnn', bias_init_with_prob=lambda p: float(-math.log((1 - p) / p)))
    mod('mmcv.cnn.bricks')
    mod('mmcv.cnn.bricks.transformer', MultiheadAttention=MultiheadAttention, FFN=FFN)
    mod('mmdet'); mod('mmdet.models'); mod('mmdet.models.utils')
    mod('mmdet.models.utils.builder', TRANSFORMER=_Registry())
    for name in ('checkpoint', 'sparsebev_transformer'):
        spec = importlib.util.spec_from_file_location('models.' + name, '%s/%s.py' % (REF, name))
        m = importlib.util.module_from_spec(spec)
        sys.modules['models.' + name] = m
        spec.loader.exec_module(m)
    return sys.modules['models.sparsebev_transformer']


def main():
    torch.set_num_threads(4)
    ref = import_reference_decoder()
    assert ref.MSMV_CUDA is False
    out = {}
    for tag, name, T, B, L in (('a', 'tiny', 2, 2, 3), ('b', 'tiny5', 3, 1, 2), ('c', 'tiny', 2, 2, 2)):
        cfg = S.layer_cfg(name, T, num_layers=L)
        sd = S.make_state_dict(cfg, seed=11)
        model = ref.SparseBEVTransformer(256, num_frames=T, num_points=cfg['num_points'], num_layers=L, num_levels=cfg['num_levels'],
                                         num_classes=10, code_size=10, pc_range=cfg['pc_range'])
        missing = model.load_state_dict({'decoder.decoder_layer.' + k: v for k, v in sd.items()}, strict=True)
        model.eval()
        feats = S.make_feats(name, T, batch=B, seed=12)
        metas = S.make_metas(name, T, batch=B)
        Q = cfg['num_query']
        n = int(np.ceil(np.sqrt(Q))) ** 2
        qb = S.init_query_bbox(n, seed=13)[:Q][None].repeat(B, 1, 1).contiguous()
        qb[..., 8:10] = 0.3 * torch.randn(B, Q, 2, generator=torch.Generator().manual_seed(14))
        qf = torch.randn(B, Q, 256, generator=torch.Generator().manual_seed(15))
        mask = None
        if tag == 'c':          # query-denoising style attention mask (sparsebev_head.py:190-207: True = this pair may not attend)
            mask = torch.rand(Q, Q, generator=torch.Generator().manual_seed(16)) < 0.3
            mask.fill_diagonal_(False)                       # no fully masked row
            out[tag + '_mask'] = mask.numpy()
        with torch.no_grad():
            cls, box = model(qb.clone(), qf.clone(), [f.clone() for f in feats], mask, copy.deepcopy(metas))
        out.update({tag + '_cfg': np.array([T, B, L]), tag + '_qb': qb.numpy(), tag + '_qf': qf.numpy(),
                    tag + '_cls': cls.numpy(), tag + '_box': box.numpy(),
                    tag + '_check': np.array([float(feats[0].double().sum()), float(sd['mixing.out_proj.weight'].double().sum())])})
        print(tag, name, 'T', T, 'B', B, 'layers', L, 'cls', tuple(cls.shape), 'box', tuple(box.shape), missing)
    np.savez(os.path.join(OUT, 'decoder.npz'), names=np.array(['tiny', 'tiny5', 'tiny']), **out)
    print('wrote', os.path.join(OUT, 'decoder.npz'))


if __name__ == '__main__':
    main()
