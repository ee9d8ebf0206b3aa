"""Eval-path mirror of the reference's `SparseBEVHead.forward` (models/sparsebev_head.py:49-117,216-220).

The reference head subclasses mmdet's DETRHead (absent in this image) and also carries the
training-only losses / query denoising, which stay with the reference.  This class reproduces what
the inference path needs -- the learned query boxes (`init_query_bbox`), the label embedding that
seeds `query_feat` (`label_enc`), the call into the transformer and the de-normalisation/reordering
of the predictions -- under the SAME parameter names, so `pts_bbox_head.*` checkpoint entries load.
With mmdet installed, keep the reference's own head and let its `transformer=dict(type='SparseBEVTransformer')`
resolve to `sparsebev_b200.transformer.SparseBEVTransformer` (registered under the same name).
"""
import math

import torch
import torch.nn as nn

from .coder import NMSFreeCoder
from .transformer import SparseBEVTransformer


class SparseBEVHead(nn.Module):
    def __init__(self, num_classes=10, in_channels=256, num_query=900, code_size=10, pc_range=None, transformer=None,
                 bbox_coder=None, **kwargs):
        super().__init__()
        ccfg = dict(bbox_coder or {})
        ccfg.pop('type', None)
        self.bbox_coder = NMSFreeCoder(**ccfg) if ccfg else None
        self.num_classes, self.embed_dims, self.num_query, self.code_size = num_classes, in_channels, num_query, code_size
        self.pc_range = list(pc_range)
        tcfg = dict(transformer or {})
        tcfg.pop('type', None)
        self.transformer = SparseBEVTransformer(**tcfg)
        self._init_layers()

    def _init_layers(self):
        self.init_query_bbox = nn.Embedding(self.num_query, 10)       # (x, y, z, w, l, h, sin, cos, vx, vy)
        self.label_enc = nn.Embedding(self.num_classes + 1, self.embed_dims - 1)
        with torch.no_grad():
            w = self.init_query_bbox.weight
            w[:, 2:3].zero_()
            w[:, 8:10].zero_()
            w[:, 5:6].fill_(1.5)
            grid = int(math.isqrt(self.num_query))
            assert grid * grid == self.num_query
            ii, jj = torch.meshgrid(torch.arange(grid), torch.arange(grid), indexing='ij')
            w[:, 0] = ((ii + 0.5) / grid).reshape(-1)
            w[:, 1] = ((jj + 0.5) / grid).reshape(-1)

    def init_weights(self):
        self.transformer.init_weights()

    @torch.no_grad()
    def forward(self, mlvl_feats, img_metas):
        B = mlvl_feats[0].shape[0]
        query_bbox = self.init_query_bbox.weight.clone()[None].repeat(B, 1, 1)
        feat = torch.cat([self.label_enc.weight[self.num_classes].repeat(self.num_query, 1),
                          torch.zeros(self.num_query, 1, device=query_bbox.device)], dim=1)
        query_feat = feat[None].repeat(B, 1, 1)
        cls_scores, bbox_preds = self.transformer(query_bbox, query_feat, mlvl_feats, attn_mask=None, img_metas=img_metas)
        pc = self.pc_range
        bbox_preds[..., 0] = bbox_preds[..., 0] * (pc[3] - pc[0]) + pc[0]
        bbox_preds[..., 1] = bbox_preds[..., 1] * (pc[4] - pc[1]) + pc[1]
        bbox_preds[..., 2] = bbox_preds[..., 2] * (pc[5] - pc[2]) + pc[2]
        bbox_preds = torch.cat([bbox_preds[..., 0:2], bbox_preds[..., 3:5], bbox_preds[..., 2:3], bbox_preds[..., 5:10]], dim=-1)
        return {'all_cls_scores': cls_scores, 'all_bbox_preds': bbox_preds, 'enc_cls_scores': None, 'enc_bbox_preds': None}

    @torch.no_grad()
    def get_bboxes(self, preds_dicts, img_metas=None, rescale=False):
        """Reference :463-482 without the mmdet3d box class: -> per sample [bboxes [n,9] (bottom-centre z), scores, labels].
        (The reference wraps `bboxes` into LiDARInstance3DBoxes(bboxes, 9).)  With utils.VERSION.name == 'v0.17.1' the
        legacy-checkpoint convention of :472-476 applies: w / l swapped, yaw -> -yaw - pi/2."""
        from .utils import VERSION
        if self.bbox_coder is None:
            raise RuntimeError('SparseBEVHead was built without bbox_coder=dict(type="NMSFreeCoder", ...)')
        out = []
        for preds in self.bbox_coder.decode(preds_dicts):
            bboxes = preds['bboxes']
            bboxes[:, 2] = bboxes[:, 2] - bboxes[:, 5] * 0.5
            if VERSION.legacy:
                w, l = bboxes[:, 3].clone(), bboxes[:, 4].clone()
                bboxes[:, 3], bboxes[:, 4] = l, w
                bboxes[:, 6] = -bboxes[:, 6] - math.pi / 2
            out.append([bboxes, preds['scores'], preds['labels']])
        return out
