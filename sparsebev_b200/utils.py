"""Mirrors of the two process-wide switches in the reference's `models/utils.py` that reach the hot path.

  VERSION   models/utils.py:320-325.  `VERSION.name = 'v0.17.1'` (set from a checkpoint's `version` key, val.py:128-129)
            selects the legacy box conventions: rotation_3d_in_axis turns the other way (models/utils.py:66-71 -> our
            sample-point kernels, option "legacy_rotation") and SparseBEVHead.get_bboxes swaps w / l and flips the yaw
            (models/sparsebev_head.py:472-476 -> sparsebev_b200/head.py).
  DUMP      models/utils.py:309-317.  When `DUMP.enabled`, the decoder writes the tensors viz_sample_points.py /
            viz_bbox_predictions.py read: per stage `sample_points_cam_stage{i}.pth` [B,T,6,Q,GP,3] and
            `sample_points_cam_valid_mask_stage{i}.pth` [B,T,6,Q,GP] (models/sparsebev_sampling.py:82-86),
            `sasa_tau_stage{i}.pth` (models/sparsebev_transformer.py:218-219), `bbox_preds_stage{i}.pth` /
            `cls_scores_stage{i}.pth` (:185-191).  A slow export path: the fused kernels never materialise these, so they are
            recomputed with a few torch ops from the kernel's inputs (transformer.py: _dump_stage).
"""
import tempfile


class DumpConfig:
    def __init__(self):
        self.enabled = False
        self.out_dir = tempfile.mkdtemp()
        self.stage_count = 0
        self.frame_count = 0


DUMP = DumpConfig()


class Version:
    """`name` behaves like the reference's plain attribute; assigning it also switches the kernels' rotation convention."""

    def __init__(self):
        self._name = 'v1.0.0'

    @property
    def name(self):
        return self._name

    @name.setter
    def name(self, value):
        self._name = value
        from . import _lib
        legacy = 1 if value == 'v0.17.1' else 0
        if _lib.get_option('legacy_rotation') != legacy:
            _lib.set_option('legacy_rotation', legacy)

    @property
    def legacy(self):
        return self._name == 'v0.17.1'


VERSION = Version()


def load_checkpoint_version(checkpoint):
    """What val.py:128-129 does after load_checkpoint: `if 'version' in checkpoint: VERSION.name = checkpoint['version']`."""
    if isinstance(checkpoint, dict) and 'version' in checkpoint:
        VERSION.name = checkpoint['version']
    return VERSION.name
