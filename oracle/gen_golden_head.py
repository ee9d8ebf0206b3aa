"""Generate tests/golden/head.npz by running the REAL reference head (eval path) -- TEST INFRASTRUCTURE.

/root/reference/models/sparsebev_head.py is executed UNMODIFIED on the CPU (imported in place, never copied):
`SparseBEVHead._init_layers` (learned query boxes on the sqrt(Q) grid, label embedding), `forward` (eval branch of
`prepare_for_dn_input`, the call into the real reference transformer of oracle/gen_golden_decoder.py, de-normalisation to
metres and the [cx, cy, w, l, cz, h, ...] reorder) and `get_bboxes` (real reference NMSFreeCoder + bottom-centre shift).

Absent third-party names are stubbed with arithmetic-free stand-ins:
  mmdet.models.dense_heads.DETRHead     nn.Module whose __init__ keeps `num_query`, builds the transformer from its cfg with the
                                        reference's own class and calls `self._init_layers()` (what mmdet 2.28.2's DETRHead does on
                                        this path; its losses / positional encoding / fc layers are never used by SparseBEVHead)
  mmdet.models.HEADS                    identity-decorator registry;  mmcv.runner.force_fp32: identity decorator factory
  mmdet.core.multi_apply / reduce_mean  unused on the eval path
  mmdet3d...build_bbox_coder            builds the REAL reference NMSFreeCoder (oracle/gen_golden_coder.py recipe)
  mmdet3d...LiDARInstance3DBoxes        keeps the raw [n, 9] tensor (`.tensor`), the form sparsebev_b200.head returns

Run in the build container only: `python oracle/gen_golden_head.py`.
"""
import copy
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
from oracle.gen_golden import OUT, REF                                   # noqa: E402
from oracle.gen_golden_decoder import import_reference_decoder, _Registry  # noqa: E402
from oracle.gen_golden_coder import import_reference_coder               # noqa: E402
from sparsebev_b200 import synthetic as S                                # noqa: E402  (input generators only)


def import_reference_head():
    ref_tr = import_reference_decoder()
    Coder = import_reference_coder()            # re-stubs models / models.bbox as namespace packages: same paths, harmless

    class DETRHead(nn.Module):
        def __init__(self, num_classes, in_channels, num_query=100, transformer=None, train_cfg=None, test_cfg=None, **kwargs):
            nn.Module.__init__(self)
            self.num_query = num_query
            tcfg = dict(transformer)
            assert tcfg.pop('type') == 'SparseBEVTransformer'
            self.transformer = ref_tr.SparseBEVTransformer(**tcfg)
            self._init_layers()

    class LiDARInstance3DBoxes:
        def __init__(self, tensor, box_dim=7):
            self.tensor, self.box_dim = tensor, box_dim

    def mod(name, **attrs):
        m = sys.modules.get(name) or types.ModuleType(name)
        m.__dict__.setdefault('__path__', [])
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    def build_bbox_coder(cfg):
        cfg = dict(cfg)
        assert cfg.pop('type') == 'NMSFreeCoder'
        return Coder(**cfg)
    mod('mmcv.runner', force_fp32=lambda *a, **k: (lambda f: f))
    mod('mmdet.core', multi_apply=None, reduce_mean=None)
    mod('mmdet.models', HEADS=_Registry())
    mod('mmdet.models.dense_heads', DETRHead=DETRHead)
    for n in ('mmdet3d', 'mmdet3d.core', 'mmdet3d.core.bbox', 'mmdet3d.core.bbox.structures'):
        mod(n)
    mod('mmdet3d.core.bbox.coders', build_bbox_coder=build_bbox_coder)
    mod('mmdet3d.core.bbox.structures.lidar_box3d', LiDARInstance3DBoxes=LiDARInstance3DBoxes)
    # models.utils (VERSION) was loaded by import_reference_decoder -> import_reference; the coder recipe replaced the
    # `models` namespace stubs, so make sure the already-loaded submodules are still reachable
    spec = importlib.util.spec_from_file_location('models.sparsebev_head', REF + '/sparsebev_head.py')
    m = importlib.util.module_from_spec(spec)
    sys.modules['models.sparsebev_head'] = m
    spec.loader.exec_module(m)
    return m


def main():
    torch.set_num_threads(4)
    ref = import_reference_head()
    name, T, B, L = 'tiny', 2, 2, 2
    cfg = S.layer_cfg(name, T, num_layers=L)
    Q = cfg['num_query']
    post = [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]
    head = ref.SparseBEVHead(num_classes=10, in_channels=256, num_query=Q, query_denoising=True, code_size=10,
                             bbox_coder=dict(type='NMSFreeCoder', pc_range=cfg['pc_range'], post_center_range=post, max_num=20,
                                             score_threshold=None, num_classes=10),
                             transformer=dict(type='SparseBEVTransformer', embed_dims=256, num_frames=T, num_points=cfg['num_points'],
                                              num_layers=L, num_levels=cfg['num_levels'], num_classes=10, code_size=10,
                                              pc_range=cfg['pc_range']))
    init_w = head.init_query_bbox.weight.detach().clone()          # the reference's own initialisation (grid / zeros / 1.5 + N(0,1))
    sd = S.make_state_dict(cfg, seed=21)
    head.transformer.load_state_dict({'decoder.decoder_layer.' + k: v for k, v in sd.items()}, strict=True)
    head.eval()
    feats = S.make_feats(name, T, batch=B, seed=22)
    metas = S.make_metas(name, T, batch=B)
    with torch.no_grad():
        outs = head([f.clone() for f in feats], copy.deepcopy(metas))
        dets = head.get_bboxes({k: (v.clone() if v is not None else None) for k, v in outs.items()}, metas)
    assert outs['enc_cls_scores'] is None and 'dn_mask_dict' not in outs
    out = dict(cfg=np.array([T, B, L, Q]), post_center_range=np.array(post, np.float32),
               init_query_bbox=init_w.numpy(), label_enc=head.label_enc.weight.detach().numpy(),
               all_cls_scores=outs['all_cls_scores'].numpy(), all_bbox_preds=outs['all_bbox_preds'].numpy(),
               check=np.array([float(feats[0].double().sum()), float(sd['mixing.out_proj.weight'].double().sum())]))
    for b, (boxes, scores, labels) in enumerate(dets):
        out['det%d_boxes' % b], out['det%d_scores' % b], out['det%d_labels' % b] = boxes.tensor.numpy(), scores.numpy(), labels.numpy()
    np.savez(os.path.join(OUT, 'head.npz'), **out)
    print('wrote', os.path.join(OUT, 'head.npz'), {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
