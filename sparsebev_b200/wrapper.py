"""Drop-in for the reference's `models/csrc/wrapper.py` (same names, argument meaning, errors).

Reference surface (paths under /root/reference):
  MSMV_CUDA                      models/csrc/wrapper.py:4-11   (flag the decoder reads, sparsebev_transformer.py:13,78)
  msmv_sampling_pytorch          models/csrc/wrapper.py:14-38
  MSMVSamplingC2345 / C23456     models/csrc/wrapper.py:41-84  (autograd Functions, positional signature kept)
  msmv_sampling                  models/csrc/wrapper.py:87-93  (dispatcher)

Differences, on purpose:
  * the CUDA path is libsparsebev_b200.so (hand-written sm_100a kernels behind a C ABI), any number of
    levels 1..5 goes to CUDA (the reference only has 4- and 5-level kernels), and kernels are enqueued on
    the CURRENT torch stream (the reference uses the legacy default stream);
  * there is no silent fallback: `msmv_sampling` on CUDA tensors raises if the library is missing or a
    launch fails (the reference prints a warning at import / printf's kernel errors and carries on).
"""
import torch
import torch.nn.functional as F

from . import _lib, ops

try:
    _lib.load()
    MSMV_CUDA = True
except (RuntimeError, OSError) as e:          # library not built: stay importable, but every CUDA call raises
    print('sparsebev_b200: CUDA library unavailable (%s); msmv_sampling on CUDA tensors will raise.' % e)
    MSMV_CUDA = False


def msmv_sampling_pytorch(mlvl_feats, sampling_locations, scale_weights):
    """Eager grid_sample formulation kept for API parity (channel-FIRST feats [B,C,N,H,W], as in the
    reference).  Not used by any sparsebev_b200 code path -- call it explicitly if you want it."""
    assert scale_weights.shape[-1] == len(mlvl_feats)
    B, C = mlvl_feats[0].shape[:2]
    _, Q, P, _ = sampling_locations.shape
    grid = (sampling_locations * 2 - 1).unsqueeze(3)                       # [B,Q,P,1,3]
    total = None
    for lvl, feat in enumerate(mlvl_feats):
        s = F.grid_sample(feat, grid, mode='bilinear', padding_mode='zeros', align_corners=True).squeeze(-1)
        s = s * scale_weights[..., lvl].reshape(B, 1, Q, P)
        total = s if total is None else total + s
    return total.permute(0, 2, 1, 3)


def _deterministic(feats):
    """Deterministic grad_feats (per-pixel segmented reduction instead of atomics) when PyTorch asks for deterministic
    algorithms (`torch.use_deterministic_algorithms(True)`) and the kernel supports the shape (C = 64)."""
    return torch.are_deterministic_algorithms_enabled() and feats[0].shape[-1] == 64


class _MSMVSamplingBase(torch.autograd.Function):
    NUM_LEVELS = 0

    @classmethod
    def _fwd(cls, ctx, *args):
        feats, sampling_locations, scale_weights = args[:cls.NUM_LEVELS], args[-2], args[-1]
        ctx.save_for_backward(*feats, sampling_locations, scale_weights)
        return ops.msmv_forward(list(feats), sampling_locations, scale_weights)

    @classmethod
    def _bwd(cls, ctx, grad_output):
        saved = ctx.saved_tensors
        feats, sampling_locations, scale_weights = saved[:cls.NUM_LEVELS], saved[-2], saved[-1]
        grad_feats, grad_loc, grad_w = ops.msmv_backward(grad_output.contiguous(), list(feats), sampling_locations,
                                                         scale_weights, deterministic=_deterministic(feats))
        return (*grad_feats, grad_loc, grad_w)


class MSMVSamplingC2345(_MSMVSamplingBase):
    NUM_LEVELS = 4

    @staticmethod
    def forward(ctx, feat_c2, feat_c3, feat_c4, feat_c5, sampling_locations, scale_weights):
        return MSMVSamplingC2345._fwd(ctx, feat_c2, feat_c3, feat_c4, feat_c5, sampling_locations, scale_weights)

    @staticmethod
    def backward(ctx, grad_output):
        return MSMVSamplingC2345._bwd(ctx, grad_output)


class MSMVSamplingC23456(_MSMVSamplingBase):
    NUM_LEVELS = 5

    @staticmethod
    def forward(ctx, feat_c2, feat_c3, feat_c4, feat_c5, feat_c6, sampling_locations, scale_weights):
        return MSMVSamplingC23456._fwd(ctx, feat_c2, feat_c3, feat_c4, feat_c5, feat_c6, sampling_locations, scale_weights)

    @staticmethod
    def backward(ctx, grad_output):
        return MSMVSamplingC23456._bwd(ctx, grad_output)


class _MSMVSamplingAnyLevels(torch.autograd.Function):
    """1, 2 or 3 levels (no counterpart kernel in the reference, which falls back to grid_sample there)."""

    @staticmethod
    def forward(ctx, sampling_locations, scale_weights, *feats):
        ctx.save_for_backward(sampling_locations, scale_weights, *feats)
        return ops.msmv_forward(list(feats), sampling_locations, scale_weights)

    @staticmethod
    def backward(ctx, grad_output):
        sampling_locations, scale_weights, *feats = ctx.saved_tensors
        grad_feats, grad_loc, grad_w = ops.msmv_backward(grad_output.contiguous(), feats, sampling_locations, scale_weights,
                                                         deterministic=_deterministic(feats))
        return (grad_loc, grad_w, *grad_feats)


def msmv_sampling(mlvl_feats, sampling_locations, scale_weights):
    """mlvl_feats: L x [B', N, H, W, C] channel-last (the layout the reference uses when MSMV_CUDA is True);
    sampling_locations [B',Q,P,3]; scale_weights [B',Q,P,L]  ->  [B',Q,C,P]."""
    if not sampling_locations.is_cuda:
        raise RuntimeError('sparsebev_b200.msmv_sampling needs CUDA tensors; there is no CPU fallback '
                           '(msmv_sampling_pytorch is the explicit eager formulation).')
    if not MSMV_CUDA:
        _lib.load()            # raises with the build hint
    if len(mlvl_feats) == 4:
        return MSMVSamplingC2345.apply(*mlvl_feats, sampling_locations, scale_weights)
    if len(mlvl_feats) == 5:
        return MSMVSamplingC23456.apply(*mlvl_feats, sampling_locations, scale_weights)
    return _MSMVSamplingAnyLevels.apply(sampling_locations, scale_weights, *mlvl_feats)
