"""Generate tests/golden/*.npz by running the REAL reference (imported in place from
/root/reference, never copied) on portable synthetic inputs -- TEST INFRASTRUCTURE.

Run in the build container only (`python oracle/gen_golden.py`); the fixtures it writes are
committed, because /root/reference does not exist on the GPU box.  Big inputs (feature maps,
mixing weights) are regenerated from `oracle.synth.hashrand` seeds by the tests; the fixtures
hold the small inputs and the reference's outputs.

What gets pinned (reference file -> fixture):
  models/csrc/wrapper.py:14-38  msmv_sampling_pytorch            -> op_small.npz, op_cfg1.npz
  models/sparsebev_sampling.py:8-24  make_sample_points          -> sampling4d.npz
  models/sparsebev_sampling.py:27-130 sampling_4d (+ the loc/w it hands the op) -> sampling4d.npz
  models/bbox/utils.py:63-77 decode_bbox, models/utils.py:87-102 inverse_sigmoid -> geometry.npz
  models/sparsebev_transformer.py:320-387 AdaptiveMixing (class body exec'd)     -> mixing.npz
"""
import importlib.util
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_torch as R          # noqa: E402
from oracle.synth import hashrand          # noqa: E402

REF = '/root/reference/models'
OUT = os.path.join(ROOT, 'tests', 'golden')


def import_reference():
    """SURVEY.md section 9 recipe: stub the packages whose __init__ needs mmcv/mmdet."""
    for name, path in [('models', REF), ('models.bbox', REF + '/bbox'), ('models.csrc', REF + '/csrc')]:
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod

    bbox_utils = load('models.bbox.utils', REF + '/bbox/utils.py')
    utils = load('models.utils', REF + '/utils.py')
    wrapper = load('models.csrc.wrapper', REF + '/csrc/wrapper.py')
    sampling = load('models.sparsebev_sampling', REF + '/sparsebev_sampling.py')
    # AdaptiveMixing: exec just that class body (the module itself needs mmcv)
    src = open(REF + '/sparsebev_transformer.py').read()
    start = src.index('class AdaptiveMixing(nn.Module):')
    ns = {'torch': torch, 'nn': torch.nn, 'F': torch.nn.functional, 'cp': None}
    exec(compile(src[start:], 'sparsebev_transformer.py[AdaptiveMixing]', 'exec'), ns)
    return bbox_utils, utils, wrapper, sampling, ns['AdaptiveMixing']


def pattern_feat(shape, seed):
    """integer-hash feature map, exact in fp32 everywhere"""
    return hashrand(shape, seed, -1.0, 1.0)


def main():
    torch.set_num_threads(4)
    os.makedirs(OUT, exist_ok=True)
    bbox_utils, utils, wrapper, sampling, AdaptiveMixing = import_reference()
    assert wrapper.MSMV_CUDA is False

    # ---- geometry helpers
    boxes = hashrand((3, 7, 10), 11, -1.5, 1.5)
    boxes[..., 0:3] = hashrand((3, 7, 3), 12, -0.1, 1.1)
    pc = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    np.savez(os.path.join(OUT, 'geometry.npz'), boxes=boxes.numpy(), pc_range=np.array(pc, np.float32),
             decoded=bbox_utils.decode_bbox(boxes, pc).numpy(),
             inv_sig=utils.inverse_sigmoid(boxes[..., 0:3]).numpy())

    # ---- the op, small multi-level multi-view case (channel-first layout for the reference)
    Bp, C, N, Q, P = 2, 64, 6, 24, 4
    hw = [(6, 10), (3, 5), (5, 7)]
    feats = [pattern_feat((Bp, C, N, h, w), 100 + i) for i, (h, w) in enumerate(hw)]
    loc = hashrand((Bp, Q, P, 3), 21, -0.15, 1.15)
    loc[..., 2] = torch.from_numpy(np.random.RandomState(5).randint(0, N, size=(Bp, Q, P))).float() / (N - 1)
    loc[0, 0, 0, :2] = torch.tensor([0.0, 0.0])      # exact corners / borders
    loc[0, 0, 1, :2] = torch.tensor([1.0, 1.0])
    loc[0, 0, 2, :2] = torch.tensor([1.0, 0.0])
    loc[0, 0, 3, :2] = torch.tensor([0.5, 0.5])
    loc[0, 1, 0, :2] = torch.tensor([-0.05, 0.3])
    loc[0, 1, 1, :2] = torch.tensor([1.04, 0.7])
    w = torch.softmax(hashrand((Bp, Q, P, len(hw)), 22, -2.0, 2.0), dim=-1)
    out = wrapper.msmv_sampling_pytorch(feats, loc, w)
    np.savez(os.path.join(OUT, 'op_small.npz'), hw=np.array(hw), feat_seeds=np.array([100, 101, 102]),
             shape=np.array([Bp, C, N, Q, P]), loc=loc.numpy(), w=w.numpy(), out=out.contiguous().numpy())

    # ---- the op, BASELINE config 1: 1 cam 256x256, 1 level, 100 queries
    feats1 = [pattern_feat((1, 64, 1, 256, 256), 300)]
    loc1 = hashrand((1, 100, 4, 3), 31, -0.1, 1.1)
    loc1[..., 2] = 0.0
    w1 = torch.ones(1, 100, 4, 1)
    out1 = wrapper.msmv_sampling_pytorch(feats1, loc1, w1)
    np.savez(os.path.join(OUT, 'op_cfg1.npz'), hw=np.array([(256, 256)]), feat_seeds=np.array([300]),
             shape=np.array([1, 64, 1, 100, 4]), loc=loc1.numpy(), w=w1.numpy(), out=out1.contiguous().numpy())

    # ---- make_sample_points + sampling_4d (T=2 frames, G=4 groups, P=4, 2 levels)
    B, Q, T, G, P, L, C = 2, 30, 2, 4, 4, 2, 64
    image_h, image_w = 64, 176
    hw4 = [(8, 22), (4, 11)]
    qb = R.init_query_bbox(36, seed=3)[:Q][None].repeat(B, 1, 1).clone()
    qb[1, :, 0:2] = hashrand((Q, 2), 41, 0.05, 0.95)
    qb[..., 3:6] = hashrand((B, Q, 3), 42, -0.5, 1.5)
    qb[..., 8:10] = hashrand((B, Q, 2), 43, -0.6, 0.6)
    offset = hashrand((B, Q, G * P, 3), 44, -0.5, 0.5)
    pts = sampling.make_sample_points(qb, offset, pc)                     # [B,Q,GP,3]
    l2i, stamps = R.camera_rig(T, image_h, image_w)
    l2i = l2i[None].repeat(B, 1, 1, 1).contiguous()
    td = R.time_diff_from_timestamps([stamps] * B)
    pts6 = pts.reshape(B, Q, 1, G, P, 3).expand(B, Q, T, G, P, 3)
    shift = qb[..., 8:10][:, :, None, :] * td[:, None, :, None]
    pts6 = torch.cat([pts6[..., 0:2] - shift[:, :, :, None, None, :], pts6[..., 2:3]], dim=-1).contiguous()
    sw = torch.softmax(hashrand((B, Q, G, 1, P, L), 45, -2, 2), dim=-1).expand(B, Q, G, T, P, L).contiguous()
    feats4 = [pattern_feat((B * T * G, C, 6, h, w_), 400 + i) for i, (h, w_) in enumerate(hw4)]
    captured = {}
    real_op = sampling.msmv_sampling

    def spy(mlvl, loc_, w_):
        captured['loc'] = loc_.clone()
        captured['w'] = w_.clone()
        return real_op(mlvl, loc_, w_)
    sampling.msmv_sampling = spy
    out4 = sampling.sampling_4d(pts6, feats4, sw, l2i, image_h, image_w)
    sampling.msmv_sampling = real_op
    np.savez(os.path.join(OUT, 'sampling4d.npz'), hw=np.array(hw4), feat_seeds=np.array([400, 401]),
             dims=np.array([B, Q, T, G, P, L, C, image_h, image_w]), pc_range=np.array(pc, np.float32),
             query_bbox=qb.numpy(), offset=offset.numpy(), points=pts.numpy(), points6=pts6.numpy(),
             lidar2img=l2i.numpy(), time_diff=td.numpy(), scale_weights=sw.numpy(),
             loc=captured['loc'].numpy(), w=captured['w'].numpy(), out=out4.contiguous().numpy())

    # ---- AdaptiveMixing (in_points = T*P = 8, G=4, C=64, out_points=128)
    Bm, Qm, Pin = 1, 12, 8
    mix = AdaptiveMixing(in_dim=256, in_points=Pin, n_groups=4, out_points=128).eval()
    sd = {
        'parameter_generator.weight': hashrand((4 * (64 * 64 + Pin * 128), 256), 501, -0.04, 0.04),
        'parameter_generator.bias': hashrand((4 * (64 * 64 + Pin * 128),), 502, -0.1, 0.1),
        'out_proj.weight': hashrand((256, 4 * 128 * 64), 503, -0.02, 0.02),
        'out_proj.bias': hashrand((256,), 504, -0.05, 0.05),
    }
    mix.load_state_dict(sd)
    x = hashrand((Bm, Qm, 4, Pin, 64), 505, -2, 2)
    query = hashrand((Bm, Qm, 256), 506, -1.5, 1.5)
    with torch.no_grad():
        ym = mix(x, query)
    np.savez(os.path.join(OUT, 'mixing.npz'), dims=np.array([Bm, Qm, 4, Pin, 64]),
             seeds=np.array([501, 502, 503, 504, 505, 506]), out=ym.numpy())

    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
