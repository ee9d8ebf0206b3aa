"""Synthetic workloads for tests and bench.py (there is no dataset / checkpoint access here).

Shapes follow the reference configs (/root/reference/configs/r50_nuimg_704x256.py:20-29,
r101_nuimg_1408x512.py, vov99_dd3d_1600x640_trainval_future.py); the distributions follow
SURVEY.md section 8(d): head-initialised query boxes, N(0, 0.02)-like "trained" weights (NOT the
zero-init of init_weights, which would make the dynamic mixing parameters query-independent), and a
nuScenes-like 6-camera pinhole rig.
"""
import math

import torch

NUM_GROUPS = 4
NUM_HEADS = 8
OUT_POINTS = 128
FFN_DIM = 512
NUM_VIEWS = 6

# name -> (image_h, image_w, level sizes, num_query, num_levels)
CONFIGS = {
    'r50_704x256': dict(image_h=256, image_w=704, levels=[(64, 176), (32, 88), (16, 44), (8, 22)], num_query=900),
    'r101_1408x512': dict(image_h=512, image_w=1408, levels=[(128, 352), (64, 176), (32, 88), (16, 44), (8, 22)], num_query=900),
    'vov99_1600x640': dict(image_h=640, image_w=1600, levels=[(160, 400), (80, 200), (40, 100), (20, 50), (10, 25)], num_query=1600),
    'tiny': dict(image_h=64, image_w=176, levels=[(8, 22), (4, 11)], num_query=36),
    'tiny5': dict(image_h=80, image_w=192, levels=[(20, 48), (10, 24), (5, 12), (3, 6), (2, 3)], num_query=40),   # 5 levels like r101 / vov99
}
PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]


def layer_cfg(name, num_frames, num_points=4, num_layers=6, num_classes=10):
    c = CONFIGS[name]
    return dict(num_frames=num_frames, num_points=num_points, num_levels=len(c['levels']), num_layers=num_layers,
                num_classes=num_classes, code_size=10, pc_range=PC_RANGE, image_h=c['image_h'], image_w=c['image_w'],
                num_query=c['num_query'], levels=c['levels'])


def make_feats(name, num_frames, batch=1, seed=0, device='cpu', embed=256, memory_format='nchw'):
    """FPN-like feature maps L x [B, T*6, 256, H, W] ~ N(0,1).  memory_format 'nhwc' returns the same logical
    tensors stored channels-last (so the decoder can read them without the regroup copy)."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    feats = []
    for (h, w) in CONFIGS[name]['levels']:
        if memory_format == 'nhwc':
            f = torch.randn(batch, num_frames * NUM_VIEWS, h, w, embed, generator=g).to(device).permute(0, 1, 4, 2, 3)
        else:
            f = torch.randn(batch, num_frames * NUM_VIEWS, embed, h, w, generator=g).to(device)
        feats.append(f)
    return feats


def make_metas(name, num_frames, batch=1):
    c = CONFIGS[name]
    l2i, stamps = camera_rig(num_frames, c['image_h'], c['image_w'])
    return [dict(img_shape=[(c['image_h'], c['image_w'], 3)] * (num_frames * NUM_VIEWS),
                 lidar2img=[m for m in l2i.numpy()], img_timestamp=list(stamps)) for _ in range(batch)]


def make_state_dict(cfg, seed=0, std=0.02, embed=256):
    """Random 'trained-like' weights (N(0,std), LayerNorm ~1/0) under the reference key names."""
    g = torch.Generator().manual_seed(seed)
    T, P, L = cfg['num_frames'], cfg['num_points'], cfg['num_levels']
    G, C = NUM_GROUPS, embed // NUM_GROUPS
    pin = T * P

    def lin(name, o, i, sd, s=std, bias_s=0.02):
        sd[name + '.weight'] = torch.randn(o, i, generator=g) * s
        sd[name + '.bias'] = torch.randn(o, generator=g) * bias_s

    def ln(name, sd):
        sd[name + '.weight'] = 1.0 + 0.1 * torch.randn(embed, generator=g)
        sd[name + '.bias'] = 0.1 * torch.randn(embed, generator=g)

    sd = {}
    lin('position_encoder.0', embed, 3, sd, s=0.5)
    ln('position_encoder.1', sd)
    lin('position_encoder.3', embed, embed, sd, s=0.06)
    ln('position_encoder.4', sd)
    sd['self_attn.attention.attn.in_proj_weight'] = torch.randn(3 * embed, embed, generator=g) * 0.06
    sd['self_attn.attention.attn.in_proj_bias'] = torch.randn(3 * embed, generator=g) * 0.02
    lin('self_attn.attention.attn.out_proj', embed, embed, sd, s=0.06)
    lin('self_attn.gen_tau', NUM_HEADS, embed, sd, s=0.02)
    sd['self_attn.gen_tau.bias'] = torch.rand(NUM_HEADS, generator=g) * 2.0
    lin('sampling.sampling_offset', G * P * 3, embed, sd, s=0.02)
    sd['sampling.sampling_offset.bias'] = torch.rand(G * P * 3, generator=g) - 0.5
    lin('sampling.scale_weights', G * P * L, embed, sd, s=0.06)
    lin('mixing.parameter_generator', G * (C * C + pin * OUT_POINTS), embed, sd, s=0.02, bias_s=0.05)
    lin('mixing.out_proj', embed, G * OUT_POINTS * C, sd, s=0.01)
    lin('ffn.layers.0.0', FFN_DIM, embed, sd, s=0.06)
    lin('ffn.layers.1', embed, FFN_DIM, sd, s=0.05)
    for n in ('norm1', 'norm2', 'norm3'):
        ln(n, sd)
    for i in (0, 3):
        lin('cls_branch.%d' % i, embed, embed, sd, s=0.06)
        ln('cls_branch.%d' % (i + 1), sd)
    lin('cls_branch.6', cfg['num_classes'], embed, sd, s=0.06)
    for i in (0, 2):
        lin('reg_branch.%d' % i, embed, embed, sd, s=0.06)
    lin('reg_branch.4', cfg.get('code_size', 10), embed, sd, s=0.02)
    return sd


def init_query_bbox(num_query, seed=0):
    """SparseBEVHead._init_layers (sparsebev_head.py:49-64): xy on a sqrt(Q) grid, z=0, h=1.5,
    v=0, everything else N(0,1) (nn.Embedding default)."""
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(num_query, 10, generator=g)
    n = int(math.isqrt(num_query))
    assert n * n == num_query
    ii, jj = torch.meshgrid(torch.arange(n), torch.arange(n), indexing='ij')
    w[:, 0] = ((ii + 0.5) / n).reshape(-1)
    w[:, 1] = ((jj + 0.5) / n).reshape(-1)
    w[:, 2] = 0.0
    w[:, 5] = 1.5
    w[:, 8:10] = 0.0
    return w


def camera_rig(num_frames, image_h, image_w, ego_speed=5.0, dt=0.5):
    """Synthetic nuScenes-like 6-camera pinhole rig -> lidar2img [T*6,4,4] fp32, timestamps [T*6].
    Order FRONT, FRONT_RIGHT, FRONT_LEFT, BACK, BACK_LEFT, BACK_RIGHT (loaders/pipelines/loading.py:54-57);
    yaw 0,-55,+55,180,+110,-110 deg; fx=fy=1266*image_w/1600; camera 1.5 m above the lidar origin;
    frame t is `dt*t` seconds in the past with the ego translated `ego_speed*dt*t` m backwards."""
    import numpy as np
    yaws = np.deg2rad([0.0, -55.0, 55.0, 180.0, 110.0, -110.0])
    f = 1266.0 * image_w / 1600.0
    K = np.array([[f, 0, image_w / 2.0, 0], [0, f, image_h / 2.0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float64)
    mats, stamps = [], []
    for t in range(num_frames):
        for n, yaw in enumerate(yaws):
            # camera axes in lidar frame: z_cam = forward (cos yaw, sin yaw, 0), x_cam = right, y_cam = down
            fwd = np.array([np.cos(yaw), np.sin(yaw), 0.0])
            right = np.array([np.sin(yaw), -np.cos(yaw), 0.0])
            down = np.array([0.0, 0.0, -1.0])
            R = np.stack([right, down, fwd])                     # lidar -> cam rotation
            cam_pos = np.array([-ego_speed * dt * t, 0.0, 1.5]) + 0.5 * fwd
            E = np.eye(4)
            E[:3, :3] = R
            E[:3, 3] = -R @ cam_pos
            mats.append(K @ E)
            stamps.append(1000.0 - dt * t - 0.004 * n)
    return torch.from_numpy(np.stack(mats).astype(np.float32)), stamps
