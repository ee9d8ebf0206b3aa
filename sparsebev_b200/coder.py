"""Post-processing mirror: NMS-free box decoding (the caller right after the decoder).

Reference: /root/reference/models/bbox/coders/nms_free_coder.py:9-110 (`NMSFreeCoder`, registry BBOX_CODERS) and
bbox/utils.py:23-45 (`denormalize_bbox`).  Same constructor keywords, `decode(preds_dicts)` / `decode_single`
signatures and result dictionaries (`bboxes` [n,9] = cx,cy,cz,w,l,h,yaw,vx,vy; `scores` [n]; `labels` [n] int64).
The work is one top-k over Q*num_classes scores and a gather of <= max_num rows -- microseconds next to the decoder --
so it stays a handful of tensor ops on whatever device the head outputs live on; nothing here calls into the CUDA library.
"""
import torch

try:                                            # register under the reference's name when mmdet is importable
    from mmdet.core.bbox.builder import BBOX_CODERS
except Exception:                               # pragma: no cover - mmdet is absent in the build container
    BBOX_CODERS = None


def denormalize_bbox(normalized_bboxes):
    """(cx, cy, log w, log l, cz, log h, sin, cos[, vx, vy]) -> (cx, cy, cz, w, l, h, yaw[, vx, vy])."""
    nb = normalized_bboxes
    yaw = torch.atan2(nb[..., 6:7], nb[..., 7:8])
    parts = [nb[..., 0:2], nb[..., 4:5], nb[..., 2:4].exp(), nb[..., 5:6].exp(), yaw]
    if nb.size(-1) > 8:
        parts.append(nb[..., 8:10])
    return torch.cat(parts, dim=-1)


class NMSFreeCoder:
    def __init__(self, pc_range, voxel_size=None, post_center_range=None, max_num=100, score_threshold=None, num_classes=10):
        self.pc_range, self.voxel_size = pc_range, voxel_size
        self.post_center_range = post_center_range
        self.max_num, self.score_threshold, self.num_classes = max_num, score_threshold, num_classes

    def encode(self):
        pass

    def decode_single(self, cls_scores, bbox_preds):
        """cls_scores [Q, num_classes] logits, bbox_preds [Q, 10] -> dict(bboxes, scores, labels)."""
        if self.post_center_range is None:
            raise NotImplementedError('Need to reorganize output as a batch, only support post_center_range is not None for now!')
        scores, flat = cls_scores.sigmoid().reshape(-1).topk(self.max_num)
        labels = flat % self.num_classes
        boxes = denormalize_bbox(bbox_preds[torch.div(flat, self.num_classes, rounding_mode='trunc')])
        limit = torch.tensor(self.post_center_range, device=scores.device)
        centre = boxes[..., :3]
        keep = (centre >= limit[:3]).all(1) & (centre <= limit[3:]).all(1)
        if self.score_threshold:                     # None and 0 both mean "no threshold", as in the reference
            keep &= scores > self.score_threshold
        return {'bboxes': boxes[keep], 'scores': scores[keep], 'labels': labels[keep]}

    def decode(self, preds_dicts):
        """Uses the LAST decoder layer: all_cls_scores [nl, B, Q, num_classes], all_bbox_preds [nl, B, Q, 10]."""
        cls, box = preds_dicts['all_cls_scores'][-1], preds_dicts['all_bbox_preds'][-1]
        return [self.decode_single(cls[b], box[b]) for b in range(cls.size(0))]


if BBOX_CODERS is not None:                     # pragma: no cover
    try:
        BBOX_CODERS.register_module(module=NMSFreeCoder, force=True)
    except Exception:
        pass
