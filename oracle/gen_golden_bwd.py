"""Generate tests/golden/op_bwd.npz: gradients of the REAL reference op -- TEST INFRASTRUCTURE.

The reference's CUDA backward (models/csrc/msmv_sampling/msmv_sampling_backward.cu) cannot run in the build
container (no GPU), but its native-PyTorch formulation of the same op, `msmv_sampling_pytorch`
(models/csrc/wrapper.py:14-38, imported in place from /root/reference, never copied), is differentiable on the CPU:
autograd through its F.grid_sample calls yields d out / d feats, d out / d (u, v) and d out / d scale_weights.  Those pin
the backward restatement in oracle/msmv_oracle.c (tests/test_oracle_golden.py::test_op_backward_matches_reference_autograd),
which in turn is what the CUDA backward kernels (atomic and deterministic) are compared with on the GPU.

Not pinned by this fixture: the gradient w.r.t. the view coordinate -- the reference CUDA kernel never writes it
(msmv_sampling_backward.cu: grad of loc[..., 2] stays 0, reproduced by our kernels), while grid_sample differentiates its
trilinear depth axis.  Sample points are kept off exact pixel centres, where bilinear interpolation is not differentiable
and the two formulations may pick different one-sided derivatives.

Run in the build container only: `python oracle/gen_golden_bwd.py`.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.gen_golden import import_reference, OUT      # noqa: E402
from oracle.synth import hashrand                        # noqa: E402


def main():
    torch.set_num_threads(4)
    _, _, wrapper, _, _ = import_reference()
    assert wrapper.MSMV_CUDA is False
    Bp, C, N, Q, P = 2, 64, 6, 14, 4
    hw = [(5, 6), (3, 4), (2, 2)]
    seeds = [400, 401, 402]
    feats = [hashrand((Bp, C, N, h, w), s, -1.0, 1.0).requires_grad_() for (h, w), s in zip(hw, seeds)]
    loc = hashrand((Bp, Q, P, 3), 41, -0.2, 1.2)            # includes the partially / fully outside bands
    loc[..., 2] = torch.from_numpy(np.random.RandomState(7).randint(0, N, size=(Bp, Q, P))).float() / (N - 1)
    for (h, w) in hw:                                       # keep every point >= 0.02 px away from a pixel centre at every level
        for axis, size in ((0, w), (1, h)):
            x = loc[..., axis] * (size - 1)
            near = (x - x.round()).abs() < 0.02
            loc[..., axis] = torch.where(near, loc[..., axis] + 0.031 / max(size - 1, 1), loc[..., axis])
    for (h, w) in hw:
        for axis, size in ((0, w), (1, h)):
            x = loc[..., axis] * (size - 1)
            assert float((x - x.round()).abs().min()) > 0.005
    loc.requires_grad_()
    wts = torch.softmax(hashrand((Bp, Q, P, len(hw)), 42, -2.0, 2.0), dim=-1).requires_grad_()
    out = wrapper.msmv_sampling_pytorch(feats, loc, wts)
    grad_out = hashrand(tuple(out.shape), 43, -1.0, 1.0)
    out.backward(grad_out)
    np.savez(os.path.join(OUT, 'op_bwd.npz'), hw=np.array(hw), feat_seeds=np.array(seeds), shape=np.array([Bp, C, N, Q, P]),
             loc=loc.detach().numpy(), w=wts.detach().numpy(), grad_out=grad_out.numpy(), out=out.detach().numpy(),
             grad_loc_uv=loc.grad[..., :2].numpy(), grad_w=wts.grad.numpy(),
             **{'grad_feat%d' % i: f.grad.permute(0, 2, 3, 4, 1).contiguous().numpy() for i, f in enumerate(feats)})   # stored channel-last
    print('wrote', os.path.join(OUT, 'op_bwd.npz'))


if __name__ == '__main__':
    main()
