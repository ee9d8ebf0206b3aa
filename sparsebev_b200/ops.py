"""Thin torch-tensor front-ends of the C ABI (device pointers + the current CUDA stream).

PyTorch here is plumbing only: allocation, stream handle, dtype/contiguity checks.  Every function
launches hand-written sm_100a kernels through libsparsebev_b200.so; nothing falls back to eager torch.
"""
import collections
import ctypes

import torch

from . import _lib

GROUPS = 4          # reference: num_groups hard-coded, sparsebev_transformer.py:123
OUT_POINTS = 128    # reference: out_points hard-coded, sparsebev_transformer.py:124
DENSE_RELU = 1
DENSE_RES_PRE_LN = 2
DENSE_REFINE = 4
DENSE_WIDE_CTA = 8


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError('%s must be a tensor' % name)
    if not t.is_cuda:
        raise RuntimeError('%s must be a CUDA tensor' % name)          # reference: AT_ASSERTM(is_cuda), msmv_sampling.cpp:113-118
    if not t.is_contiguous():
        raise RuntimeError('%s tensor has to be contiguous' % name)    # reference: msmv_sampling.cpp:106-111
    if t.dtype != dtype:
        raise RuntimeError('%s must be %s (got %s)' % (name, dtype, t.dtype))
    return t


def _p(t):
    return None if t is None else t.data_ptr()


def _levels(mlvl_feats):
    L = len(mlvl_feats)
    hw = []
    for i, f in enumerate(mlvl_feats):
        _chk(f, 'value[%d]' % i)
        if f.dim() != 5:
            raise RuntimeError('value[%d] must be [B, N, H, W, C]' % i)
        hw += [f.shape[2], f.shape[3]]
    return L, _lib.ptr_array([f.data_ptr() for f in mlvl_feats]), _lib.i32_array(hw)


def msmv_forward(mlvl_feats, sampling_locations, scale_weights):
    """feats L x [B',N,H,W,C] channel-last, loc [B',Q,P,3], w [B',Q,P,L] -> [B',Q,C,P]."""
    lib = _lib.load()
    L, fptr, hw = _levels(mlvl_feats)
    loc = _chk(sampling_locations, 'sampling_loc')
    w = _chk(scale_weights, 'attn_weight')
    Bp, N, _, _, C = mlvl_feats[0].shape
    _, Q, P, _ = loc.shape
    if w.shape[-1] != L or tuple(w.shape[:3]) != (Bp, Q, P) or loc.shape[0] != Bp or loc.shape[3] != 3:
        raise RuntimeError('shape mismatch between value, sampling_loc and attn_weight')
    out = torch.empty(Bp, Q, C, P, device=loc.device, dtype=torch.float32)
    with torch.cuda.device(loc.device):
        _lib.check(lib.sbev_msmv_fwd(fptr, hw, L, loc.data_ptr(), w.data_ptr(), Bp, N, C, Q, P, out.data_ptr(), _stream()),
                   'sbev_msmv_fwd')
    return out


def msmv_backward(grad_output, mlvl_feats, sampling_locations, scale_weights, deterministic=False):
    """deterministic: grad_feats by the per-pixel segmented reduction (sbev_msmv_bwd_det: no floating-point atomics,
    bit-identical between runs) instead of vector atomics."""
    lib = _lib.load()
    L, fptr, hw = _levels(mlvl_feats)
    loc = _chk(sampling_locations, 'sampling_loc')
    w = _chk(scale_weights, 'attn_weight')
    go = _chk(grad_output, 'grad_output')
    Bp, N, _, _, C = mlvl_feats[0].shape
    _, Q, P, _ = loc.shape
    grad_feats = [torch.empty_like(f) for f in mlvl_feats]
    grad_loc = torch.empty_like(loc)
    grad_w = torch.empty_like(w)
    if deterministic:
        nbytes = lib.sbev_msmv_bwd_det_workspace(hw, L, Bp, N, Q, P)
        if nbytes < 0:
            raise RuntimeError('sbev_msmv_bwd_det_workspace: invalid sizes')
        work = torch.empty(nbytes // 4 + 1, dtype=torch.int32, device=loc.device)
        with torch.cuda.device(loc.device):
            _lib.check(lib.sbev_msmv_bwd_det(go.data_ptr(), fptr, hw, L, loc.data_ptr(), w.data_ptr(), Bp, N, C, Q, P,
                                             _lib.ptr_array([g.data_ptr() for g in grad_feats]), grad_loc.data_ptr(),
                                             grad_w.data_ptr(), work.data_ptr(), work.numel() * 4, _stream()), 'sbev_msmv_bwd_det')
        return grad_feats, grad_loc, grad_w
    with torch.cuda.device(loc.device):
        _lib.check(lib.sbev_msmv_bwd(go.data_ptr(), fptr, hw, L, loc.data_ptr(), w.data_ptr(), Bp, N, C, Q, P,
                                     _lib.ptr_array([g.data_ptr() for g in grad_feats]), grad_loc.data_ptr(),
                                     grad_w.data_ptr(), _stream()), 'sbev_msmv_bwd')
    return grad_feats, grad_loc, grad_w


def msmv_indices(level_hw, sampling_locations, num_views):
    lib = _lib.load()
    loc = _chk(sampling_locations, 'sampling_loc')
    Bp, Q, P, _ = loc.shape
    L = len(level_hw)
    hw = _lib.i32_array([int(v) for pair in level_hw for v in pair])
    view = torch.empty(Bp, Q, P, dtype=torch.int32, device=loc.device)
    y0 = torch.empty(Bp, Q, P, L, dtype=torch.int32, device=loc.device)
    x0 = torch.empty_like(y0)
    inside = torch.empty_like(y0)
    with torch.cuda.device(loc.device):
        _lib.check(lib.sbev_msmv_indices(hw, L, loc.data_ptr(), Bp, num_views, Q, P, view.data_ptr(), y0.data_ptr(),
                                         x0.data_ptr(), inside.data_ptr(), _stream()), 'sbev_msmv_indices')
    return view, y0, x0, inside


def sampling4d_fused(mlvl_feats, points, velocity, time_diff, lidar2img, scale_w, image_h, image_w,
                     num_frames, num_views=6, eps=1e-5, layout='grouped', return_loc=False, out=None, frame_window=None,
                     scatter_ptrs=None, owner_ptrs=None, q_per_rank=0):
    """Fused motion-warp + projection + view pick + gather.

    frame_window (t0, t1): the feature maps hold only frames [t0, t1) of the num_frames (frame-sharded decoder); the
    result is then [B,Q,G,(t1-t0)*P,C].  time_diff / lidar2img always cover all frames.
    scatter_ptrs: device addresses of full-size [B,Q,G,T*P,C] buffers (this GPU's and its peers', e.g. symmetric memory);
    the window's rows are stored into every one of them at their frame offset and no tensor is returned.
    owner_ptrs + q_per_rank (B == 1): device addresses of every rank's [q_per_rank,G,T*P,C] buffer; query q's rows are stored
    only into owner_ptrs[q // q_per_rank] (query- and frame-sharded decoder); no tensor is returned.

    layout 'grouped': feats L x [B*T*G, N, H, W, C] (the reference's regrouped op layout);
    layout 'nhwc'   : feats L x [B, T*N, H, W, G*C] (un-regrouped, channels-last FPN output).
    points [B,Q,G*P,3], velocity [B,Q,2] (or the whole query_bbox [B,Q,10], read in place), time_diff [B,T], lidar2img [B,T*N,4,4], scale_w [B,Q,G,P,L]
    -> [B,Q,G,T*P,C] (+ loc [B*T*G,Q,P,3] when return_loc)."""
    lib = _lib.load()
    L = len(mlvl_feats)
    pts = _chk(points, 'points')
    if velocity.dim() == 3 and velocity.shape[-1] == 10:          # the whole query_bbox [B,Q,10]: read vx, vy in place
        vel, vel_ptr, ld_vel = _chk(velocity, 'query_bbox'), velocity.data_ptr() + 8 * 4, 10
    else:
        vel = _chk(velocity, 'velocity')
        vel_ptr, ld_vel = vel.data_ptr(), 2
    td = _chk(time_diff, 'time_diff')
    l2i = _chk(lidar2img, 'lidar2img')
    sw = _chk(scale_w, 'scale_w')
    B, Q, GP, _ = pts.shape
    T, N, G = num_frames, num_views, GROUPS
    t0, t1 = frame_window if frame_window is not None else (0, T)
    Tl = t1 - t0
    if not (0 <= t0 <= t1 <= T):
        raise ValueError('frame_window %r outside [0, %d]' % (frame_window, T))
    P = GP // G
    hw, s_bt, s_g, s_v, s_px = [], [], [], [], []
    C = 0
    for i, f in enumerate(mlvl_feats):
        _chk(f, 'value[%d]' % i)
        if layout == 'grouped':
            BTG, Nf, H, W, C = f.shape
            if BTG != B * Tl * G or Nf != N:
                raise RuntimeError('value[%d] has shape %s, expected [%d,%d,H,W,C]' % (i, tuple(f.shape), B * Tl * G, N))
            s_px.append(C); s_v.append(H * W * C); s_g.append(N * H * W * C); s_bt.append(G * N * H * W * C)
        elif layout == 'nhwc':
            Bf, TN, H, W, GC = f.shape
            C = GC // G
            if Bf != B or TN != Tl * N:
                raise RuntimeError('value[%d] has shape %s, expected [%d,%d,H,W,G*C]' % (i, tuple(f.shape), B, Tl * N))
            s_px.append(GC); s_v.append(H * W * GC); s_g.append(C); s_bt.append(N * H * W * GC)
        else:
            raise ValueError('unknown layout %r' % layout)
        hw += [H, W]
    if tuple(sw.shape) != (B, Q, G, P, L):
        raise RuntimeError('scale_w must be [B,Q,G,P,L]=%s, got %s' % ((B, Q, G, P, L), tuple(sw.shape)))
    if tuple(td.shape) != (B, T) or tuple(l2i.shape) != (B, T * N, 4, 4) or tuple(vel.shape) != (B, Q, ld_vel):
        raise RuntimeError('time_diff / lidar2img / velocity shape mismatch')
    loc = torch.empty(B * Tl * G, Q, P, 3, device=pts.device, dtype=torch.float32) if return_loc else None
    if owner_ptrs is not None:
        if B != 1 or return_loc:
            raise RuntimeError('the owner form of the gather needs B == 1 and cannot return loc')
        with torch.cuda.device(pts.device):
            _lib.check(lib.sbev_sampling4d_owner_fwd(
                _lib.ptr_array([f.data_ptr() for f in mlvl_feats]), _lib.i32_array(hw), L,
                _lib.i64_array(s_bt), _lib.i64_array(s_g), _lib.i64_array(s_v), _lib.i64_array(s_px),
                pts.data_ptr(), vel_ptr, ld_vel, td.data_ptr(), l2i.data_ptr(), sw.data_ptr(),
                T, t0, Tl, G, N, C, Q, P, float(image_h), float(image_w), float(eps),
                _lib.ptr_array([int(p) for p in owner_ptrs]), len(owner_ptrs), int(q_per_rank), _stream()), 'sbev_sampling4d_owner_fwd')
        return None
    if scatter_ptrs is not None:
        with torch.cuda.device(pts.device):
            _lib.check(lib.sbev_sampling4d_scatter_fwd(
                _lib.ptr_array([f.data_ptr() for f in mlvl_feats]), _lib.i32_array(hw), L,
                _lib.i64_array(s_bt), _lib.i64_array(s_g), _lib.i64_array(s_v), _lib.i64_array(s_px),
                pts.data_ptr(), vel_ptr, ld_vel, td.data_ptr(), l2i.data_ptr(), sw.data_ptr(),
                B, T, t0, Tl, G, N, C, Q, P, float(image_h), float(image_w), float(eps),
                _lib.ptr_array([int(p) for p in scatter_ptrs]), len(scatter_ptrs), _p(loc), _stream()), 'sbev_sampling4d_scatter_fwd')
        return (None, loc) if return_loc else None
    if out is None:
        out = torch.empty(B, Q, G, Tl * P, C, device=pts.device, dtype=torch.float32)
    with torch.cuda.device(pts.device):
        _lib.check(lib.sbev_sampling4d_window_fwd(
            _lib.ptr_array([f.data_ptr() for f in mlvl_feats]), _lib.i32_array(hw), L,
            _lib.i64_array(s_bt), _lib.i64_array(s_g), _lib.i64_array(s_v), _lib.i64_array(s_px),
            pts.data_ptr(), vel_ptr, ld_vel, td.data_ptr(), l2i.data_ptr(), sw.data_ptr(),
            B, T, t0, Tl, G, N, C, Q, P, float(image_h), float(image_w), float(eps),
            out.data_ptr(), _p(loc), _stream()), 'sbev_sampling4d_window_fwd')
    return (out, loc) if return_loc else out


class DenseWeight:
    """Device-side cache of one (or several, concatenated along the output dim) nn.Linear weights in the layout
    the dense kernels want: Wt[K][ldw] = W^T zero-padded to a multiple of 4 columns (+ the concatenated bias).
    Rebuilt when any parameter changes (data_ptr / version / device)."""

    def __init__(self):
        self.key = None
        self.wt = None
        self.bias = None
        self.ldw = 0
        self.n = 0
        self.w_hi = self.w_lo = None
        self.w_pack = None
        self.kpad = 0

    def get(self, weight, *more_weights):
        return self.get_with_bias([weight] + list(more_weights), None)[:2]

    def get_with_bias(self, weights, biases):
        key = tuple((w.data_ptr(), w._version, tuple(w.shape), w.device) for w in weights)
        if biases is not None:
            key += tuple((b.data_ptr(), b._version) for b in biases if b is not None)
        if key != self.key:
            K = weights[0].shape[1]
            N = sum(w.shape[0] for w in weights)
            ldw = (N + 3) // 4 * 4
            wt = torch.zeros(K, ldw, device=weights[0].device, dtype=torch.float32)
            wt[:, :N] = torch.cat([w.detach() for w in weights], dim=0).t()
            self.bias = None
            if biases is not None and any(b is not None for b in biases):
                self.bias = torch.cat([b.detach() if b is not None else torch.zeros(w.shape[0], device=w.device)
                                       for w, b in zip(weights, biases)]).contiguous()
            # tensor-core operand: bf16 (hi, lo) split in the nn.Linear layout [N][Kpad], K zero-padded to a multiple of 64
            kpad = (K + 63) // 64 * 64
            wpad = torch.zeros(N, kpad, device=weights[0].device, dtype=torch.float32)
            wpad[:, :K] = torch.cat([w.detach() for w in weights], dim=0)
            self.w_hi, self.w_lo = split_bf16(wpad) if wpad.is_cuda else (None, None)
            self.w_pack = pack_weight_tiles(self.w_hi, self.w_lo) if wpad.is_cuda else None
            self.kpad = kpad
            self.wt, self.ldw, self.n, self.key = wt, ldw, N, key
        return self.wt, self.ldw, self.bias


def pack_weight_tiles(w_hi, w_lo):
    """bf16 (hi, lo) weights [N][Kpad] (Kpad % 64 == 0) -> the pre-tiled streaming copy of sbev_dense_layer.W_pack:
    [ceil(N/128)][Kpad/64][hi|lo][128][64] bf16, rows zero-padded to a multiple of 128, 16-byte chunks of every 128-byte
    row XOR-swizzled by (row & 7).  Built once per weight version; every pipeline stage of the chain kernels becomes one
    contiguous 32 KB bulk copy."""
    N, Kp = w_hi.shape
    nb, kc = (N + 127) // 128, Kp // 64
    r = torch.arange(128, device=w_hi.device)
    src_chunk = torch.arange(8, device=w_hi.device)[None, :] ^ (r[:, None] & 7)          # stored position p of row r holds chunk p ^ (r & 7)

    def one(w):
        wp = torch.zeros(nb * 128, Kp, device=w.device, dtype=w.dtype)
        wp[:N] = w
        t = wp.view(nb, 128, kc, 8, 8).permute(0, 2, 1, 3, 4)                            # [nb, kc, row, chunk, elem]
        return t[:, :, r[:, None], src_chunk, :]                                         # [nb, kc, row, position, elem]
    return torch.stack([one(w_hi), one(w_lo)], dim=2).contiguous()                       # [nb, kc, 2, 128, 8, 8]


class _ChainEntry(tuple):
    """(sbev_dense_layer struct, tensors to keep alive) + the bf16 (hi, lo) weights the entry was built from."""
    w_hi = w_lo = None


def chain_layer(wt, ldw, K, N, bias=None, ln=None, residual=None, relu=False, res_pre_ln=False, refine=False, y=None, ldy=None,
                w_hi=None, w_lo=None, kpad=0, y_hi=None, y_lo=None, w_pack=None, wide_cta=False):
    """One entry of a dense chain (see dense_chain).  w_hi / w_lo (bf16 [N][kpad]) enable the tensor-core path; w_pack
    (pack_weight_tiles) its bulk-copy weight stream."""
    flags = (DENSE_RELU if relu else 0) | (DENSE_RES_PRE_LN if res_pre_ln else 0) | (DENSE_REFINE if refine else 0) | (DENSE_WIDE_CTA if wide_cta else 0)
    keep = [t for t in (wt, bias, residual, y, w_hi, w_lo, y_hi, y_lo, w_pack) if t is not None] + ([ln.weight, ln.bias] if ln is not None else [])
    e = _ChainEntry((_lib.DenseLayer(_p(wt), ldw, K, N, _p(bias), _p(ln.weight) if ln is not None else None,
                                     _p(ln.bias) if ln is not None else None, _p(residual), flags, _p(y),
                                     (ldy if ldy is not None else N) if (y is not None or y_hi is not None) else 0, _p(w_hi), _p(w_lo), kpad,
                                     _p(y_hi), _p(y_lo), _p(w_pack)), keep))
    e.w_hi, e.w_lo = w_hi, w_lo
    return e


WS_CLUSTER = 8                       # CTAs per cluster of the weights-stationary chain kernel (csrc/dense_ws.cu)
WS_SMEM_CAP = 226 * 1024


def ws_slice_width(N):
    """Output features per CTA of a layer in the weights-stationary chain: ceil(N / 8) rounded up to the MMA's 8-feature tile."""
    return ((N + WS_CLUSTER - 1) // WS_CLUSTER + 7) // 8 * 8


def pack_ws_blob(weights):
    """[(w_hi, w_lo)] (bf16 [N_i][Kpad_i], Kpad % 64 == 0) of a chain -> (blob uint8 [8][stride], stride): the per-CTA weight slices
    of sbev_dense_chain_ws_fwd -- rank c, layers back to back, each [Kpad/64][hi|lo][SW rows][64 k] with row j = feature c * SW + j
    (zero beyond N) and the 16-byte chunks of every 128-byte row XOR-swizzled by (j & 7)."""
    dev = weights[0][0].device
    parts = []
    for w_hi, w_lo in weights:
        N, Kp = w_hi.shape
        SW, kc = ws_slice_width(N), Kp // 64
        j = torch.arange(SW, device=dev)
        src_chunk = torch.arange(8, device=dev)[None, :] ^ (j[:, None] & 7)            # stored position p of row j holds chunk p ^ (j & 7)

        def one(w):
            wp = torch.zeros(WS_CLUSTER * SW, Kp, device=dev, dtype=w.dtype)
            wp[:N] = w
            t = wp.view(WS_CLUSTER, SW, kc, 8, 8).permute(0, 2, 1, 3, 4)                # [c, kc, row, chunk, elem]
            return t[:, :, j[:, None], src_chunk, :]                                    # [c, kc, row, position, elem]
        both = torch.stack([one(w_hi), one(w_lo)], dim=2).contiguous()                  # [c, kc, 2, SW, 8, 8]
        parts.append(both.view(WS_CLUSTER, -1).view(torch.uint8))
    blob = torch.cat(parts, dim=1)
    stride = (blob.shape[1] + 127) // 128 * 128
    out = torch.zeros(WS_CLUSTER, stride, device=dev, dtype=torch.uint8)
    out[:, :blob.shape[1]] = blob
    return out, stride


_ws_blobs = collections.OrderedDict()        # (id(w_hi), ...) -> (blob, stride, [w_hi, ...] kept alive so the ids stay unique)


def _ws_blob(layers):
    key = tuple(id(l.w_hi) for l in layers)
    hit = _ws_blobs.get(key)
    if hit is None:
        blob, stride = pack_ws_blob([(l.w_hi, l.w_lo) for l in layers])
        hit = (blob, stride, [l.w_hi for l in layers])
        _ws_blobs[key] = hit
        while len(_ws_blobs) > 64:
            _ws_blobs.popitem(last=False)
    else:
        _ws_blobs.move_to_end(key)
    return hit[0], hit[1]


def ws_eligible(M, layers, reduce_k0=None):
    """Can the weights-stationary cluster kernel run this chain (the limits listed at sbev_dense_chain_ws_fwd), and does option
    "dense_ws" (1 = every such chain, 2 = only chains of <= 256 rows) send it there?"""
    mode = _lib.get_option('dense_ws')
    if mode == 0 or (mode == 2 and M > 256) or M == 0 or not (1 <= len(layers) <= 6):
        return False
    if any(getattr(l, 'w_hi', None) is None or not l.w_hi.is_cuda for l in layers):
        return False
    if reduce_k0 is not None and reduce_k0 != 256:
        return False
    if reduce_k0 is not None and not layers[-1][0].ln_w:
        return False
    blob_bytes, kmax, swx = 0, 64, (8 if reduce_k0 is not None else 0)
    for i, l in enumerate(layers):
        d = l[0]
        last = i + 1 == len(layers)
        exchange = not (last and not d.ln_w)
        if d.Kpad > 512 or d.Kpad % 64:
            return False
        if exchange:
            if d.N > 512 or d.N % 4 or (d.flags & DENSE_REFINE) or ((d.y or d.y_hi) and d.ldy % 4):
                return False
            if any((p or 0) % 16 for p in (d.ln_w, d.ln_b, d.residual, d.y)) or any((p or 0) % 8 for p in (d.y_hi, d.y_lo)):
                return False
            swx = max(swx, ws_slice_width(d.N))
        blob_bytes += (d.Kpad // 64) * 2 * ws_slice_width(d.N) * 128
        kmax = max(kmax, d.Kpad)
    group_bytes = (2 * 16 * (kmax + 8) * 2 + (8 * 16 * swx + 2 * 16 * swx) * 4 + 127) // 128 * 128
    return (blob_bytes + 1023) // 1024 * 1024 + group_bytes + 1024 <= WS_SMEM_CAP


def dense_chain(x, ldx, M, layers, refine_proposal=None, refine_time_diff=None, refine_Q=0, refine_T=0):
    """Run 1..6 Linear(+bias)(+residual)(+LayerNorm)(+ReLU) layers in ONE kernel: the weights-stationary cluster kernel
    (sbev_dense_chain_ws_fwd) when it can express the chain and option "dense_ws" selects it, else the row-group kernel that
    streams the weights (sbev_dense_chain_fwd).  `layers` = list of chain_layer(...) results; outputs are the `y` tensors given to them."""
    lib = _lib.load()
    _chk(x, 'x')
    arr = (_lib.DenseLayer * len(layers))(*[l[0] for l in layers])
    with torch.cuda.device(x.device):
        if ws_eligible(M, layers):
            blob, stride = _ws_blob(layers)
            _lib.check(lib.sbev_dense_chain_ws_fwd(x.data_ptr(), ldx, M, len(layers), arr, blob.data_ptr(), stride,
                                                   _p(refine_proposal), _p(refine_time_diff), refine_Q, refine_T, _stream()),
                       'sbev_dense_chain_ws_fwd')
            return
        _lib.check(lib.sbev_dense_chain_fwd(x.data_ptr(), ldx, M, len(layers), arr, _p(refine_proposal), _p(refine_time_diff),
                                            refine_Q, refine_T, _stream()), 'sbev_dense_chain_fwd')


def dense_chain_points(x, ldx, M, layers, query_bbox, pc_range, GP, L, off_col, log_col):
    """dense_chain whose last layer's row holds [GP*3 sampling offsets @ off_col | GP*L scale logits @ log_col]: also returns
    the sample points [B,Q,GP,3] and softmaxed scale weights [B,Q,GP,L] (sbev_dense_chain_points_fwd, one launch)."""
    lib = _lib.load()
    _chk(x, 'x')
    qb = _chk(query_bbox, 'query_bbox')
    B, Q, _ = qb.shape
    pts = torch.empty(B, Q, GP, 3, device=qb.device, dtype=torch.float32)
    sw = torch.empty(B, Q, GP, L, device=qb.device, dtype=torch.float32)
    arr = (_lib.DenseLayer * len(layers))(*[l[0] for l in layers])
    with torch.cuda.device(x.device):
        _lib.check(lib.sbev_dense_chain_points_fwd(x.data_ptr(), ldx, M, len(layers), arr, qb.data_ptr(),
                                                   _lib.f32_array([float(v) for v in pc_range]), GP, L, off_col, log_col,
                                                   pts.data_ptr(), sw.data_ptr(), _stream()), 'sbev_dense_chain_points_fwd')
    return pts, sw


def dense_chain_reduce(partial, bias, residual, ln_w, ln_b, x_out, layers, refine_proposal=None, refine_time_diff=None, refine_Q=0, refine_T=0):
    """dense_chain whose input rows are LN(sum_z partial[z] + bias + residual) (sbev_dense_chain_reduce_fwd): the split-K
    reduce + norm of the preceding GEMM runs in the chain's prologue.  partial [S,M,K0]; x_out [M,K0] receives the input rows."""
    lib = _lib.load()
    partial = _chk(partial, 'partial')
    S, M, K0 = partial.shape
    _chk(x_out, 'x_out')
    arr = (_lib.DenseLayer * len(layers))(*[l[0] for l in layers])
    with torch.cuda.device(partial.device):
        if ws_eligible(M, layers, reduce_k0=K0) and all(t is None or t.data_ptr() % 16 == 0 for t in (partial, bias, residual, ln_w, ln_b, x_out)):
            blob, stride = _ws_blob(layers)
            _lib.check(lib.sbev_dense_chain_ws_reduce_fwd(partial.data_ptr(), S, _p(bias), _p(residual), _p(ln_w), _p(ln_b), x_out.data_ptr(),
                                                          M, len(layers), arr, blob.data_ptr(), stride,
                                                          _p(refine_proposal), _p(refine_time_diff), refine_Q, refine_T, _stream()),
                       'sbev_dense_chain_ws_reduce_fwd')
            return
        _lib.check(lib.sbev_dense_chain_reduce_fwd(partial.data_ptr(), S, _p(bias), _p(residual), _p(ln_w), _p(ln_b), x_out.data_ptr(),
                                                   M, len(layers), arr, _p(refine_proposal), _p(refine_time_diff), refine_Q, refine_T,
                                                   _stream()), 'sbev_dense_chain_reduce_fwd')


def dense(x, wt, ldw, n_out, bias=None, ln_w=None, ln_b=None, residual=None, relu=False, res_pre_ln=False, out=None, k=None):
    """y[M,N] = epilogue(x[M,K] @ W^T) with W given pre-transposed (see DenseWeight)."""
    lib = _lib.load()
    if k is None:
        x = _chk(x, 'x')
        M, K = x.shape
        ldx = K
    else:                       # use the first k columns of a wider row-major matrix without copying
        _chk(x, 'x')
        M, ldx = x.shape
        K = k
    if out is None:
        out = torch.empty(M, n_out, device=x.device, dtype=torch.float32)
    flags = (DENSE_RELU if relu else 0) | (DENSE_RES_PRE_LN if res_pre_ln else 0)
    with torch.cuda.device(x.device):
        _lib.check(lib.sbev_dense_fwd(x.data_ptr(), ldx, wt.data_ptr(), ldw, _p(bias), _p(ln_w), _p(ln_b), _p(residual),
                                      M, K, n_out, flags, out.data_ptr(), _stream()), 'sbev_dense_fwd')
    return out


def sample_points(query_bbox, offset, scale_logits, pc_range, num_levels, num_points_total=None, ld_off=None, ld_log=None, out=None):
    """query_bbox [B,Q,10]; offset rows of GP*3 floats, scale_logits rows of GP*L floats (either plain [B,Q,GP*3] /
    [B,Q,GP*L] tensors, or column blocks of a wider matrix given with explicit row strides ld_off / ld_log)
    -> points [B,Q,GP,3], scale_w [B,Q,GP,L] (GP = G*P, group-major)."""
    lib = _lib.load()
    qb = _chk(query_bbox, 'query_bbox')
    B, Q, _ = qb.shape
    L = num_levels
    if ld_off is None:
        off = _chk(offset, 'offset')
        lg = _chk(scale_logits, 'scale_logits')
        GP = off.shape[-1] // 3
        ld_off, ld_log = GP * 3, GP * L
    else:
        off, lg, GP = offset, scale_logits, num_points_total
    if out is not None:                  # caller-provided (e.g. rows of a symmetric-memory buffer)
        pts, sw = out
        if pts.numel() != B * Q * GP * 3 or sw.numel() != B * Q * GP * L:
            raise RuntimeError('sample_points: out buffers have the wrong size')
        _chk(pts, 'points'); _chk(sw, 'scale_w')
    else:
        pts = torch.empty(B, Q, GP, 3, device=qb.device, dtype=torch.float32)
        sw = torch.empty(B, Q, GP, L, device=qb.device, dtype=torch.float32)
    with torch.cuda.device(qb.device):
        _lib.check(lib.sbev_sample_points_fwd(qb.data_ptr(), off.data_ptr(), ld_off, lg.data_ptr(), ld_log,
                                              _lib.f32_array([float(v) for v in pc_range]), B * Q, GP, L,
                                              pts.data_ptr(), sw.data_ptr(), _stream()), 'sbev_sample_points_fwd')
    return pts, sw


def sasa(qkv, query_bbox, tau, pc_range, num_heads=8, dn_mask=None, ld_qkv=None, ld_tau=None, embed_dims=None):
    """qkv [B,Q,3D] and tau [B,Q,H] (or column blocks of one wider [B*Q, ld] matrix, with explicit row strides)
    + query_bbox [B,Q,10] -> attention output [B,Q,D] (before out_proj)."""
    lib = _lib.load()
    qb = _chk(query_bbox, 'query_bbox')
    B, Q, _ = qb.shape
    if ld_qkv is None:
        qkv = _chk(qkv, 'qkv')
        tau = _chk(tau, 'tau')
        D = qkv.shape[-1] // 3
        ld_qkv, ld_tau = 3 * D, num_heads
    else:
        D = embed_dims
    m = None
    if dn_mask is not None:
        m = _chk(dn_mask.to(torch.uint8).contiguous(), 'dn_mask', torch.uint8)
    out = torch.empty(B, Q, D, device=qb.device, dtype=torch.float32)
    with torch.cuda.device(qb.device):
        _lib.check(lib.sbev_sasa_fwd(qkv.data_ptr(), ld_qkv, qb.data_ptr(), tau.data_ptr(), ld_tau, _p(m),
                                     _lib.f32_array([float(v) for v in pc_range]), B, Q, num_heads, D,
                                     out.data_ptr(), _stream()), 'sbev_sasa_fwd')
    return out


def sasa_split(qkvt, query_bbox, pc_range, num_heads, embed_dims, dn_mask=None, split=None, q_range=None, out=None):
    """Tensor-core attention core on the concatenated in_proj|gen_tau output `qkvt` [B*Q, 3D+H] (fp32): splits it once
    into bf16 (hi, lo) and runs the warp-pipelined kernel.  -> [B,Q,D] (heads concatenated, before out_proj).
    q_range (qa, qb): only those queries attend (to all Q keys) and only their rows of the result are written."""
    lib = _lib.load()
    qkvt = _chk(qkvt, 'qkvt')
    qb = _chk(query_bbox, 'query_bbox')
    B, Q, _ = qb.shape
    D, H = embed_dims, num_heads
    ld = qkvt.shape[1]
    hi, lo = split if split is not None else split_bf16(qkvt)
    m = None
    if dn_mask is not None:
        m = _chk(dn_mask.to(torch.uint8).contiguous(), 'dn_mask', torch.uint8)
    if out is None:
        out = torch.empty(B, Q, D, device=qb.device, dtype=torch.float32)
    qa, qe = (0, Q) if q_range is None else q_range
    with torch.cuda.device(qb.device):
        _lib.check(lib.sbev_sasa_split_range_fwd(hi.data_ptr(), lo.data_ptr(), ld, qb.data_ptr(), qkvt.data_ptr() + 3 * D * 4, ld, _p(m),
                                                 _lib.f32_array([float(v) for v in pc_range]), B, Q, H, D, int(qa), int(qe), out.data_ptr(), _stream()),
                   'sbev_sasa_split_range_fwd')
    return out


def split_bf16(x, need_lo=True, out=None):
    """fp32 -> (hi, lo) bf16 with x ~= hi + lo.  `out` = optional (hi, lo) buffers."""
    lib = _lib.load()
    x = _chk(x, 'x')
    if out is not None:
        hi, lo = out
    else:
        hi = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
        lo = torch.empty_like(hi) if need_lo else None
    with torch.cuda.device(x.device):
        _lib.check(lib.sbev_split_bf16(x.data_ptr(), ctypes.c_int64(x.numel()), hi.data_ptr(), _p(lo), _stream()),
                   'sbev_split_bf16')
    return hi, lo


def gemm_bf16_tn(a_list, b_list, M, N, K, bias=None, split_k=1, out=None):
    """C[M,N] = sum_s A_s[M,K] @ B_s[N,K]^T (+bias) on tcgen05; returns [split_k, M, N] fp32 partials
    (a plain [M,N] when split_k == 1)."""
    lib = _lib.load()
    for t in list(a_list) + list(b_list):
        _chk(t, 'gemm operand', torch.bfloat16)
    dev = a_list[0].device
    if out is None:
        out = torch.empty((split_k, M, N) if split_k > 1 else (M, N), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(lib.sbev_gemm_bf16_tn(_lib.ptr_array([t.data_ptr() for t in a_list]),
                                         _lib.ptr_array([t.data_ptr() for t in b_list]), len(a_list),
                                         _p(bias), M, N, K, split_k, out.data_ptr(), _stream()), 'sbev_gemm_bf16_tn')
    return out


def gemm_bf16_tn_split(a_hi, a_lo, b_hi, b_lo, M, N, K, bias=None, out=None):
    """bf16x3 GEMM whose result leaves as a bf16 (hi, lo) pair [M,N] each (C ~= hi + lo)."""
    lib = _lib.load()
    for t in (a_hi, a_lo, b_hi, b_lo):
        _chk(t, 'gemm operand', torch.bfloat16)
    if out is None:
        out = (torch.empty(M, N, device=a_hi.device, dtype=torch.bfloat16), torch.empty(M, N, device=a_hi.device, dtype=torch.bfloat16))
    with torch.cuda.device(a_hi.device):
        _lib.check(lib.sbev_gemm_bf16_tn_split(a_hi.data_ptr(), a_lo.data_ptr(), b_hi.data_ptr(), b_lo.data_ptr(), _p(bias), M, N, K,
                                               out[0].data_ptr(), out[1].data_ptr(), _stream()), 'sbev_gemm_bf16_tn_split')
    return out


def mix_presplit(params_hi, params_lo, x, want_f32=False, want_split=True, out=None):
    """mix() with the dynamic parameters given as the bf16 (hi, lo) pair of gemm_bf16_tn_split (in_points must be 32).
    out = optional (hi, lo) bf16 buffers [BQ, G*Pout*C] to write into."""
    lib = _lib.load()
    _chk(params_hi, 'params_hi', torch.bfloat16)
    _chk(params_lo, 'params_lo', torch.bfloat16)
    x = _chk(x, 'x')
    BQ, G, Pin, C = x.shape
    n = G * OUT_POINTS * C
    if out is not None:
        hi, lo = out
        _chk(hi, 'y_hi', torch.bfloat16); _chk(lo, 'y_lo', torch.bfloat16)
        if hi.numel() != BQ * n or lo.numel() != BQ * n:
            raise RuntimeError('mix_presplit: out buffers must hold [BQ, G*Pout*C]')
    else:
        hi = torch.empty(BQ, n, device=x.device, dtype=torch.bfloat16) if want_split else None
        lo = torch.empty_like(hi) if want_split else None
    yf = torch.empty(BQ, n, device=x.device, dtype=torch.float32) if want_f32 else None
    with torch.cuda.device(x.device):
        _lib.check(lib.sbev_mix_presplit_fwd(params_hi.data_ptr(), params_lo.data_ptr(), x.data_ptr(), BQ, G, Pin, OUT_POINTS, C,
                                             _p(hi), _p(lo), _p(yf), _stream()), 'sbev_mix_presplit_fwd')
    return hi, lo, yf


def mix(params, x, want_f32=False, want_split=True):
    """params [BQ, G*(C*C+Pout*Pin)], x [BQ,G,Pin,C] -> (y_hi, y_lo) bf16 [BQ, G*Pout*C] (and/or fp32 y)."""
    lib = _lib.load()
    params = _chk(params, 'params')
    x = _chk(x, 'x')
    BQ, G, Pin, C = x.shape
    n = G * OUT_POINTS * C
    hi = torch.empty(BQ, n, device=x.device, dtype=torch.bfloat16) if want_split else None
    lo = torch.empty_like(hi) if want_split else None
    yf = torch.empty(BQ, n, device=x.device, dtype=torch.float32) if want_f32 else None
    with torch.cuda.device(x.device):
        _lib.check(lib.sbev_mix_fwd(params.data_ptr(), x.data_ptr(), BQ, G, Pin, OUT_POINTS, C,
                                    _p(hi), _p(lo), _p(yf), _stream()), 'sbev_mix_fwd')
    return hi, lo, yf


def reduce_ln(partial, bias=None, residual=None, ln_w=None, ln_b=None):
    """[S,M,N] (or [M,N]) partials -> LN(sum + bias + residual) [M,N]."""
    lib = _lib.load()
    partial = _chk(partial, 'partial')
    if partial.dim() == 2:
        partial = partial[None]
    S, M, N = partial.shape
    out = torch.empty(M, N, device=partial.device, dtype=torch.float32)
    with torch.cuda.device(partial.device):
        _lib.check(lib.sbev_reduce_ln_fwd(partial.data_ptr(), S, _p(bias), _p(residual), _p(ln_w), _p(ln_b), M, N,
                                          out.data_ptr(), _stream()), 'sbev_reduce_ln_fwd')
    return out


def refine_bbox(proposal, delta, time_diff):
    """proposal/delta [B,Q,code], time_diff [B,T] -> refined boxes [B,Q,code]."""
    lib = _lib.load()
    proposal = _chk(proposal, 'proposal')
    delta = _chk(delta, 'delta')
    td = _chk(time_diff, 'time_diff')
    B, Q, code = proposal.shape
    out = torch.empty_like(proposal)
    with torch.cuda.device(proposal.device):
        _lib.check(lib.sbev_refine_bbox_fwd(proposal.data_ptr(), delta.data_ptr(), td.data_ptr(), B, Q, td.shape[1], code,
                                            out.data_ptr(), _stream()), 'sbev_refine_bbox_fwd')
    return out


def peer_exchange(segments, n_peers, rank, flag_ptrs, ctl_ptr, device):
    """One kernel: copy `segments` = [(src_ptr, [dst_ptr per rank], nbytes), ...] to every peer, then an all-ranks barrier
    (sbev_peer_exchange).  No segments = pure barrier."""
    lib = _lib.load()
    arr = (_lib.PeerSegment * max(1, len(segments)))()
    for i, (src, dsts, nbytes) in enumerate(segments):
        arr[i].src = src
        for w, d in enumerate(dsts):
            arr[i].dst[w] = d
        arr[i].bytes = nbytes
    with torch.cuda.device(device):
        _lib.check(lib.sbev_peer_exchange(arr, len(segments), n_peers, rank, _lib.ptr_array([int(p) for p in flag_ptrs]), ctl_ptr, _stream()),
                   'sbev_peer_exchange')


# ------------------------------------------------------------------------------------------------
# Backbone convolutions: NHWC bf16 activations, tcgen05 implicit GEMM.
def cast_bf16(x):
    """fp32 CUDA tensor -> bf16 tensor of the same shape (our kernel; round to nearest even)."""
    lib = _lib.load()
    _chk(x, 'x')
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.check(lib.sbev_cast_bf16(x.data_ptr(), x.numel(), y.data_ptr(), _stream()), 'sbev_cast_bf16')
    return y


def conv_out_size(H, W, k, stride, pad):
    return (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1


def conv2d_nhwc(x, w, shift, scale=None, stride=1, pad=0, residual=None, relu=False, out_f32=False):
    """x NHWC bf16 [N,H,W,Cin], w bf16 [Cout,KH,KW,Cin], shift / scale fp32 [Cout], residual NHWC bf16 [N,rH,rW,Cout]
    (same size as the output, or smaller = nearest-upsampled) -> NHWC [N,Ho,Wo,Cout] bf16 | fp32."""
    lib = _lib.load()
    _chk(x, 'x', torch.bfloat16); _chk(w, 'weight', torch.bfloat16); _chk(shift, 'shift')
    if scale is not None:
        _chk(scale, 'scale')
    if x.dim() != 4 or w.dim() != 4 or w.shape[3] != x.shape[3]:
        raise RuntimeError('conv2d_nhwc: x must be [N,H,W,Cin] and weight [Cout,KH,KW,Cin]')
    N, H, W, Cin = x.shape
    Cout, KH, KW, _ = w.shape
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    rH = rW = 0
    if residual is not None:
        _chk(residual, 'residual', torch.bfloat16)
        if residual.dim() != 4 or residual.shape[0] != N or residual.shape[3] != Cout:
            raise RuntimeError('conv2d_nhwc: residual must be [N,rH,rW,Cout]')
        rH, rW = residual.shape[1], residual.shape[2]
    out = torch.empty(N, Ho, Wo, Cout, device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.check(lib.sbev_conv2d_nhwc_fwd(x.data_ptr(), N, H, W, Cin, w.data_ptr(), Cout, KH, KW, stride, pad,
                                            _p(scale), shift.data_ptr(), _p(residual), rH, rW, int(relu),
                                            out.data_ptr(), int(out_f32), _stream()), 'sbev_conv2d_nhwc_fwd')
    return out


def stem_conv(img, w, scale, shift):
    """img NCHW fp32 [N,3,H,W], w fp32 [k,k,3,64] (k = 7: ResNet stem, k = 3: VoVNet stem_1) -> relu(bn(conv kxk / 2, pad k//2)) as
    NHWC bf16 [N,Ho,Wo,64]."""
    lib = _lib.load()
    _chk(img, 'img'); _chk(w, 'weight'); _chk(scale, 'scale'); _chk(shift, 'shift')
    k = w.shape[0]
    if img.dim() != 4 or img.shape[1] != 3 or tuple(w.shape) != (k, k, 3, 64) or k not in (3, 7):
        raise RuntimeError('stem_conv: img must be [N,3,H,W] and weight [k,k,3,64], k = 3 or 7')
    N, _, H, W = img.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    out = torch.empty(N, Ho, Wo, 64, device=img.device, dtype=torch.bfloat16)
    with torch.cuda.device(img.device):
        _lib.check(lib.sbev_stem_conv_k_fwd(img.data_ptr(), N, H, W, w.data_ptr(), k, scale.data_ptr(), shift.data_ptr(),
                                            out.data_ptr(), _stream()), 'sbev_stem_conv_k_fwd')
    return out


def maxpool3x3s2_nhwc(x):
    lib = _lib.load()
    _chk(x, 'x', torch.bfloat16)
    N, H, W, C = x.shape
    out = torch.empty(N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C, device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.check(lib.sbev_maxpool3x3s2_nhwc_fwd(x.data_ptr(), N, H, W, C, out.data_ptr(), _stream()), 'sbev_maxpool3x3s2_nhwc_fwd')
    return out


def maxpool3x3s2_ex_nhwc(x, pad=0, ceil_mode=True):
    """3x3 stride-2 max pool, NHWC bf16, torch.nn.MaxPool2d(3, 2, padding=pad, ceil_mode=ceil_mode) semantics."""
    lib = _lib.load()
    _chk(x, 'x', torch.bfloat16)
    N, H, W, C = x.shape

    def osz(s):
        num = s + 2 * pad - 3
        o = ((num + 1) // 2 if ceil_mode else num // 2) + 1
        return o - 1 if (ceil_mode and (o - 1) * 2 >= s + pad) else o
    out = torch.empty(N, osz(H), osz(W), C, device=x.device, dtype=torch.bfloat16)
    with torch.cuda.device(x.device):
        _lib.check(lib.sbev_maxpool3x3s2_ex_nhwc_fwd(x.data_ptr(), N, H, W, C, int(pad), int(bool(ceil_mode)), out.data_ptr(), _stream()),
                   'sbev_maxpool3x3s2_ex_nhwc_fwd')
    return out


def ese_nhwc(x, fc_weight, fc_bias, identity=None):
    """Effective squeeze-excitation (+ optional OSA identity): x NHWC bf16 [N,H,W,C], fc_weight fp32 [C,C], fc_bias fp32 [C]
    -> x * hsigmoid(fc(mean_hw x)) (+ identity), NHWC bf16."""
    lib = _lib.load()
    _chk(x, 'x', torch.bfloat16); _chk(fc_weight, 'fc_weight'); _chk(fc_bias, 'fc_bias')
    N, H, W, C = x.shape
    if tuple(fc_weight.shape) != (C, C) or tuple(fc_bias.shape) != (C,):
        raise RuntimeError('ese_nhwc: fc_weight must be [C,C] and fc_bias [C]')
    if identity is not None:
        _chk(identity, 'identity', torch.bfloat16)
        if identity.shape != x.shape:
            raise RuntimeError('ese_nhwc: identity must have the shape of x')
    ws = torch.empty(lib.sbev_ese_workspace_floats(N, H, W, C), device=x.device, dtype=torch.float32)
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _lib.check(lib.sbev_ese_nhwc_fwd(x.data_ptr(), N, H, W, C, fc_weight.data_ptr(), fc_bias.data_ptr(), _p(identity), ws.data_ptr(),
                                         out.data_ptr(), _stream()), 'sbev_ese_nhwc_fwd')
    _lib.launch_count += 2           # three kernels of ours per call
    return out


def subsample2_nhwc(x):
    lib = _lib.load()
    _chk(x, 'x')
    N, H, W, C = x.shape
    out = torch.empty(N, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(lib.sbev_subsample2_nhwc_fwd(x.data_ptr(), N, H, W, C, out.data_ptr(), _stream()), 'sbev_subsample2_nhwc_fwd')
    return out
