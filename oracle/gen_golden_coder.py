"""Golden vectors for the post-processing row: the REAL reference NMSFreeCoder
(/root/reference/models/bbox/coders/nms_free_coder.py:37-110 + bbox/utils.py:23-45) run on seeded inputs.

TEST INFRASTRUCTURE -- runs only in the build container (imports /root/reference); the vectors travel as
tests/golden/coder.npz.  mmdet is absent, so its two names the file imports (BaseBBoxCoder, BBOX_CODERS) are stubbed with
an empty base class and an identity decorator -- neither contributes arithmetic.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.synth import hashrand   # noqa: E402

REF = '/root/reference/models'
OUT = os.path.join(ROOT, 'tests', 'golden')


def import_reference_coder():
    for name, path in [('models', REF), ('models.bbox', REF + '/bbox'), ('models.bbox.coders', REF + '/bbox/coders')]:
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m

    class _Registry:
        def register_module(self, *a, **k):
            return lambda cls: cls
    for name in ('mmdet', 'mmdet.core', 'mmdet.core.bbox', 'mmdet.core.bbox.builder'):
        sys.modules[name] = types.ModuleType(name)
    sys.modules['mmdet.core.bbox'].BaseBBoxCoder = object
    sys.modules['mmdet.core.bbox.builder'].BBOX_CODERS = _Registry()

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod
    load('models.bbox.utils', REF + '/bbox/utils.py')
    return load('models.bbox.coders.nms_free_coder', REF + '/bbox/coders/nms_free_coder.py').NMSFreeCoder


def main():
    Coder = import_reference_coder()
    nl, B, Q, C = 2, 2, 300, 10
    cls = hashrand((nl, B, Q, C), 901, -6.0, 3.0)
    box = hashrand((nl, B, Q, 10), 902, -1.0, 1.0)
    box[..., 0:2] = hashrand((nl, B, Q, 2), 903, -70.0, 70.0)      # metres, some beyond post_center_range
    box[..., 4] = hashrand((nl, B, Q), 904, -12.0, 12.0)
    post = [-61.2, -61.2, -10.0, 61.2, 61.2, 10.0]
    out = {'cls': cls.numpy(), 'box': box.numpy(), 'post_center_range': np.array(post, dtype=np.float32)}
    for tag, kw in [('a', dict(max_num=100, score_threshold=None)), ('b', dict(max_num=37, score_threshold=0.05))]:
        coder = Coder(pc_range=[-51.2, -51.2, -5.0, 51.2, 51.2, 3.0], post_center_range=post, num_classes=C, **kw)
        res = coder.decode({'all_cls_scores': cls, 'all_bbox_preds': box})
        for b, r in enumerate(res):
            out['%s%d_bboxes' % (tag, b)] = r['bboxes'].numpy()
            out['%s%d_scores' % (tag, b)] = r['scores'].numpy()
            out['%s%d_labels' % (tag, b)] = r['labels'].numpy()
    np.savez(os.path.join(OUT, 'coder.npz'), **out)
    print('coder.npz', os.path.getsize(os.path.join(OUT, 'coder.npz')), {k: v.shape for k, v in out.items() if k[0] in 'ab'})


if __name__ == '__main__':
    main()
