"""ctypes front-end of oracle/msmv_oracle.c -- TEST INFRASTRUCTURE (see that file's header)."""
import ctypes
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'msmv_oracle.c')
SO = os.path.join(HERE, '_build', 'libmsmv_oracle.so')
_lib = None


def build():
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(SRC):
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-shared', '-fPIC', SRC, '-o', SO, '-lm'])
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _fp(t):
    return ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_float))


def _ptr_array(ts):
    return (ctypes.POINTER(ctypes.c_float) * len(ts))(*[_fp(t) for t in ts])


def _dims(feats_cl, loc):
    Bp, N, _, _, C = feats_cl[0].shape
    _, Q, P, _ = loc.shape
    hw = np.array([[f.shape[2], f.shape[3]] for f in feats_cl], dtype=np.int32).reshape(-1)
    return Bp, N, C, Q, P, hw


def fwd(feats_cl, loc, w):
    feats_cl = [f.contiguous().float() for f in feats_cl]
    loc, w = loc.contiguous().float(), w.contiguous().float()
    Bp, N, C, Q, P, hw = _dims(feats_cl, loc)
    out = torch.empty(Bp, Q, C, P)
    lib().msmv_oracle_fwd(_ptr_array(feats_cl), hw.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), len(feats_cl),
                          _fp(loc), _fp(w), Bp, N, C, Q, P, _fp(out))
    return out


def indices(level_hw, loc, num_views):
    loc = loc.contiguous().float()
    Bp, Q, P, _ = loc.shape
    L = len(level_hw)
    hw = np.array(level_hw, dtype=np.int32).reshape(-1)
    view = torch.empty(Bp, Q, P, dtype=torch.int32)
    y0 = torch.empty(Bp, Q, P, L, dtype=torch.int32)
    x0 = torch.empty_like(y0)
    ins = torch.empty_like(y0)
    ip = lambda t: ctypes.cast(t.data_ptr(), ctypes.POINTER(ctypes.c_int32))
    lib().msmv_oracle_indices(hw.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), L, _fp(loc), Bp, num_views, Q, P,
                              ip(view), ip(y0), ip(x0), ip(ins))
    return view, y0, x0, ins


def bwd(grad_out, feats_cl, loc, w):
    feats_cl = [f.contiguous().float() for f in feats_cl]
    loc, w, grad_out = loc.contiguous().float(), w.contiguous().float(), grad_out.contiguous().float()
    Bp, N, C, Q, P, hw = _dims(feats_cl, loc)
    gf = [torch.empty_like(f) for f in feats_cl]
    gl, gw = torch.empty_like(loc), torch.empty_like(w)
    lib().msmv_oracle_bwd(_fp(grad_out), _ptr_array(feats_cl), hw.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                          len(feats_cl), _fp(loc), _fp(w), Bp, N, C, Q, P, _ptr_array(gf), _fp(gl), _fp(gw))
    return gf, gl, gw
