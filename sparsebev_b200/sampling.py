"""Drop-in for the reference's `models/sparsebev_sampling.py` (make_sample_points, sampling_4d).

Reference: /root/reference/models/sparsebev_sampling.py:8-24 and :27-130.  Same signatures and
tensor layouts; the work happens in two CUDA kernels (sbev_sample_points_fwd for the box decode /
rotation, sbev_sampling4d_fwd for projection + view pick + gather) instead of ~40 eager launches.
"""
import torch

from . import ops
from .wrapper import msmv_sampling, msmv_sampling_pytorch  # noqa: F401  (re-exported like the reference module)

NUM_VIEWS = 6     # reference hard-codes N = 6 (sparsebev_sampling.py:45)


def _forward_only(what, *tensors):
    """These two functions launch the forward kernels directly (no autograd Function): used in a training graph they would
    silently cut the gradient to the sampling offsets, the scale weights and every backbone / FPN feature while the loss
    still back-propagates through mixing, FFN and the heads.  Fail loudly instead; the differentiable drop-in is level 1
    of INTEGRATION.md (`sparsebev_b200.wrapper.msmv_sampling` under the reference's own sampling_4d)."""
    if torch.is_grad_enabled():
        for t in tensors:
            ts = t if isinstance(t, (list, tuple)) else [t]
            if any(torch.is_tensor(x) and x.requires_grad for x in ts):
                raise RuntimeError('sparsebev_b200.sampling.%s is forward-only (no backward kernel for the fused front-end): call it under '
                                   'torch.no_grad(), or keep the reference\'s %s and swap only models.csrc.wrapper for '
                                   'sparsebev_b200.wrapper, whose msmv_sampling has a backward (INTEGRATION.md, level 1)' % (what, what))


def make_sample_points(query_bbox, offset, pc_range):
    """query_bbox [B,Q,10], offset [B,Q,GP,3] -> [B,Q,GP,3] lidar-frame points."""
    _forward_only('make_sample_points', query_bbox, offset)
    B, Q, GP, _ = offset.shape
    logits = torch.zeros(B, Q, GP, device=offset.device, dtype=torch.float32)     # 1 level of logits, result unused
    pts, _ = ops.sample_points(query_bbox.contiguous().float(), offset.reshape(B, Q, GP * 3).contiguous().float(),
                               logits, pc_range, num_levels=1)
    return pts


def sampling_4d(sample_points, mlvl_feats, scale_weights, lidar2img, image_h, image_w, eps=1e-5):
    """sample_points [B,Q,T,G,P,3] (already motion-warped, as the reference passes them);
    mlvl_feats L x [B*T*G, N, H, W, C] channel-last; scale_weights [B,Q,G,T,P,L]; lidar2img [B,T*N,4,4]
    -> [B,Q,G,T*P,C].

    The fused kernel wants the un-warped points + velocity; to keep THIS signature (warped points per
    frame) we hand it per-frame points through a zero velocity: points of frame t are read from
    sample_points[:, :, t]."""
    _forward_only('sampling_4d', sample_points, mlvl_feats, scale_weights)
    B, Q, T, G, P, _ = sample_points.shape
    L = scale_weights.shape[-1]
    zero_v = torch.zeros(B, Q, 2, device=sample_points.device, dtype=torch.float32)
    zero_t = torch.zeros(B, 1, device=sample_points.device, dtype=torch.float32)
    # scale weights may differ per t in this general signature -> one launch per frame slice
    outs = []
    sw = scale_weights.reshape(B, Q, G, T, P, L)
    C = mlvl_feats[0].shape[-1]
    for t in range(T):
        feats_t = [f.reshape(B, T, G, *f.shape[1:])[:, t].reshape(B * G, *f.shape[1:]) for f in mlvl_feats]
        if not all(f.is_contiguous() for f in feats_t):
            feats_t = [f.contiguous() for f in feats_t]
        # weight row for loc slice (b,t,g) is that of (g',t') = divmod(t*G+g, T): gather it explicitly
        idx = torch.arange(G, device=sw.device) + t * G
        gsel, tsel = idx // T, idx % T
        sw_t = sw[:, :, gsel, tsel]                                     # [B,Q,G,P,L]
        out_t = ops.sampling4d_fused(feats_t, sample_points[:, :, t].reshape(B, Q, G * P, 3).contiguous(),
                                     zero_v, zero_t, lidar2img.reshape(B, T, NUM_VIEWS, 4, 4)[:, t].contiguous(),
                                     sw_t.contiguous(), image_h, image_w, num_frames=1, num_views=NUM_VIEWS, eps=eps)
        outs.append(out_t)                                              # [B,Q,G,P,C]
    return torch.stack(outs, dim=3).reshape(B, Q, G, T * P, C)
