"""Same-name stand-in for the reference's pybind extension `_msmv_sampling_cuda`
(/root/reference/models/csrc/msmv_sampling/msmv_sampling.cpp:362-369), for code that imports the extension module
directly instead of going through `wrapper.py`:

    _ms_deform_attn_cuda_c2345_forward(f2, f3, f4, f5, sampling_loc, attn_weight) -> Tensor [B', Q, C, P]
    _ms_deform_attn_cuda_c2345_backward(grad_output, f2, f3, f4, f5, sampling_loc, attn_weight)
        -> [grad_f2, grad_f3, grad_f4, grad_f5, grad_sampling_loc, grad_attn_weight]
    _ms_deform_attn_cuda_c23456_forward / _backward: the same with a fifth level f6.

Every call goes to libsparsebev_b200.so through the C ABI (sbev_msmv_fwd / sbev_msmv_bwd) on the current torch stream;
there is no fallback.  Argument checks follow the reference's AT_ASSERTM messages (`... must be contiguous`,
`num_point exceed limits`), raised as RuntimeError.
"""
from . import ops


def _ms_deform_attn_cuda_c2345_forward(feat_c2, feat_c3, feat_c4, feat_c5, sampling_loc, attn_weight):
    return ops.msmv_forward([feat_c2, feat_c3, feat_c4, feat_c5], sampling_loc, attn_weight)


def _ms_deform_attn_cuda_c2345_backward(grad_output, feat_c2, feat_c3, feat_c4, feat_c5, sampling_loc, attn_weight):
    grad_feats, grad_loc, grad_w = ops.msmv_backward(grad_output, [feat_c2, feat_c3, feat_c4, feat_c5], sampling_loc, attn_weight)
    return [*grad_feats, grad_loc, grad_w]


def _ms_deform_attn_cuda_c23456_forward(feat_c2, feat_c3, feat_c4, feat_c5, feat_c6, sampling_loc, attn_weight):
    return ops.msmv_forward([feat_c2, feat_c3, feat_c4, feat_c5, feat_c6], sampling_loc, attn_weight)


def _ms_deform_attn_cuda_c23456_backward(grad_output, feat_c2, feat_c3, feat_c4, feat_c5, feat_c6, sampling_loc, attn_weight):
    grad_feats, grad_loc, grad_w = ops.msmv_backward(grad_output, [feat_c2, feat_c3, feat_c4, feat_c5, feat_c6], sampling_loc, attn_weight)
    return [*grad_feats, grad_loc, grad_w]
