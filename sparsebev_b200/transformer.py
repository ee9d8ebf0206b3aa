"""Drop-in for the reference's `models/sparsebev_transformer.py`, inference path, on B200 kernels.

Same class names, constructor kwargs, call signatures, side effects and state-dict keys as
/root/reference/models/sparsebev_transformer.py, so a reference checkpoint loads unchanged and
`SparseBEVHead` can call `self.transformer(query_bbox, query_feat, mlvl_feats, attn_mask=..., img_metas=...)`
(reference: models/sparsebev_head.py:77-83) without modification:

    SparseBEVTransformer            :16-38     (registered as TRANSFORMER when mmdet is importable)
    SparseBEVTransformerDecoder     :41-101    (shared-weight layer looped num_layers times)
    SparseBEVTransformerDecoderLayer:104-193
    SparseBEVSelfAttention          :196-248   (`.attention.attn` = nn.MultiheadAttention parameters, `.gen_tau`)
    SparseBEVSampling               :251-317
    AdaptiveMixing                  :320-387

The nn.Modules only HOLD parameters (so names/shapes match the checkpoint); every forward runs the
hand-written sm_100a kernels in libsparsebev_b200.so through `ops` -- 11 launches per layer, capturable in one
CUDA graph (bench.py does).  Forward only: this
is the eval path (dropout = identity, no activation checkpointing); training the decoder through these
modules is out of scope (the op-level autograd Functions in `wrapper.py` do have a backward).
"""
import collections
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from .utils import DUMP

try:                                                     # mmcv / mmdet are optional (absent in the build image)
    from mmcv.runner import BaseModule
except Exception:                                        # pragma: no cover
    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

try:
    from mmdet.models.utils.builder import TRANSFORMER
except Exception:                                        # pragma: no cover
    TRANSFORMER = None

NUM_VIEWS = 6     # hard-coded in the reference (sparsebev_transformer.py:61,75)


def decode_bbox(bboxes, pc_range=None):
    """Torch mirror of models/bbox/utils.py:63-77 (used by the DUMP export only; the kernels decode in registers)."""
    xyz = bboxes[..., 0:3].clone()
    wlh = bboxes[..., 3:6].exp()
    rot = torch.atan2(bboxes[..., 6:7], bboxes[..., 7:8])
    if pc_range is not None:
        for i in range(3):
            xyz[..., i] = xyz[..., i] * (pc_range[3 + i] - pc_range[i]) + pc_range[i]
    if bboxes.shape[-1] > 8:
        return torch.cat([xyz, wlh, rot, bboxes[..., 8:10].clone()], dim=-1)
    return torch.cat([xyz, wlh, rot], dim=-1)


def projected_sample_points(points, velocity, time_diff, lidar2img, image_h, image_w, eps=1e-5, num_views=NUM_VIEWS):
    """Slow export path of what the fused gather keeps in registers: all T x N projections of the (motion-warped) sample
    points, in the form models/sparsebev_sampling.py:82-86 dumps for viz_sample_points.py.
    points [B,Q,GP,3], velocity [B,Q,2], time_diff [B,T], lidar2img [B,T*N,4,4]
    -> (cam [B,T,N,Q,GP,3] = (u / image_w, v / image_h, max(depth, eps)), valid_mask [B,T,N,Q,GP] float)."""
    B, Q, GP, _ = points.shape
    T = time_diff.shape[1]
    p = points[:, :, None].expand(B, Q, T, GP, 3)
    xy = p[..., 0:2] - (velocity[:, :, None, :] * time_diff[:, None, :, None])[:, :, :, None, :]       # sparsebev_transformer.py:286-295
    ph = torch.cat([xy, p[..., 2:3], torch.ones_like(p[..., :1])], dim=-1)                             # [B,Q,T,GP,4]
    l2i = lidar2img.reshape(B, T, num_views, 4, 4)
    cam = torch.einsum('btnij,bqtpj->btnqpi', l2i, ph)                                                  # [B,T,N,Q,GP,4]
    homo = cam[..., 2:3]
    homo_nonzero = torch.maximum(homo, torch.zeros_like(homo) + eps)
    uv = cam[..., 0:2] / homo_nonzero
    uv = torch.stack([uv[..., 0] / image_w, uv[..., 1] / image_h], dim=-1)
    valid = ((homo > eps) & (uv[..., 1:2] > 0.0) & (uv[..., 1:2] < 1.0) & (uv[..., 0:1] > 0.0) & (uv[..., 0:1] < 1.0)).squeeze(-1).float()
    return torch.cat([uv, homo_nonzero], dim=-1), valid


class _Dense:
    """One nn.Linear (or several sharing the input, concatenated along the output dim) + optional LayerNorm,
    bound to the cached device layout the dense kernels want.  `layer(...)` builds a chain entry."""

    def __init__(self, linear, ln=None, extra=()):
        self.linears, self.ln, self.cache = [linear] + list(extra), ln, ops.DenseWeight()
        self.in_features = linear.in_features
        self.out_features = sum(l.out_features for l in self.linears)

    def layer(self, relu=False, residual=None, res_pre_ln=False, refine=False, y=None, ldy=None, y_hi=None, y_lo=None, wide_cta=False):
        wt, ldw, bias = self.cache.get_with_bias([l.weight for l in self.linears], [l.bias for l in self.linears])
        c = self.cache
        return ops.chain_layer(wt, ldw, self.in_features, self.out_features, bias=bias, ln=self.ln, residual=residual,
                               relu=relu, res_pre_ln=res_pre_ln, refine=refine, y=y, ldy=ldy, w_hi=c.w_hi, w_lo=c.w_lo, kpad=c.kpad,
                               y_hi=y_hi, y_lo=y_lo, w_pack=c.w_pack, wide_cta=wide_cta)

    def __call__(self, x, relu=False, residual=None, res_pre_ln=False, k=None):
        M = x.shape[0]
        y = torch.empty(M, self.out_features, device=x.device, dtype=torch.float32)
        ops.dense_chain(x, x.shape[1], M, [self.layer(relu=relu, residual=residual, res_pre_ln=res_pre_ln, y=y)])
        return y


class _SplitWeight:
    """bf16 (hi, lo) split of an nn.Linear weight [N,K] for the bf16x3 tcgen05 GEMM; cached per version."""

    def __init__(self):
        self.key, self.hi, self.lo = None, None, None

    def get(self, weight):
        key = (weight.data_ptr(), weight._version, tuple(weight.shape), weight.device)
        if key != self.key:
            self.hi, self.lo = ops.split_bf16(weight.detach().contiguous())
            self.key = key
        return self.hi, self.lo


class AdaptiveMixing(nn.Module):
    """Adaptive Mixing (reference :320-387).  parameter_generator and out_proj run as bf16x3 tcgen05 GEMMs
    (fp32-grade accuracy, `precision='bf16x3'`) or single-pass bf16 (`precision='bf16'`)."""

    def __init__(self, in_dim, in_points, n_groups=1, query_dim=None, out_dim=None, out_points=None):
        super().__init__()
        out_dim = out_dim if out_dim is not None else in_dim
        out_points = out_points if out_points is not None else in_points
        query_dim = query_dim if query_dim is not None else in_dim
        self.query_dim, self.in_dim, self.in_points, self.n_groups = query_dim, in_dim, in_points, n_groups
        self.out_dim, self.out_points = out_dim, out_points
        self.eff_in_dim, self.eff_out_dim = in_dim // n_groups, out_dim // n_groups
        self.m_parameters = self.eff_in_dim * self.eff_out_dim
        self.s_parameters = self.in_points * self.out_points
        self.total_parameters = self.m_parameters + self.s_parameters
        self.parameter_generator = nn.Linear(self.query_dim, self.n_groups * self.total_parameters)
        self.out_proj = nn.Linear(self.eff_out_dim * self.out_points * self.n_groups, self.query_dim)
        self.act = nn.ReLU(inplace=True)
        self.precision = 'bf16x3'
        self.split_k = 18            # 8 M-tiles x 18 K-slices = 144 CTAs on 148 SMs
        # parameters leave the GEMM as bf16 (hi, lo) and reach the mix kernel by TMA (needs in_points == 32, bf16x3)
        self.tma_params = True
        self._pg, self._op = _SplitWeight(), _SplitWeight()

    @torch.no_grad()
    def init_weights(self):
        nn.init.zeros_(self.parameter_generator.weight)

    def alloc_params(self, M, device):
        """Buffers of the parameter-generation stage (allocated by the caller BEFORE it forks a side stream)."""
        n_par = self.n_groups * self.total_parameters
        buf = dict(q_hi=torch.empty(M, self.query_dim, device=device, dtype=torch.bfloat16),
                   q_lo=torch.empty(M, self.query_dim, device=device, dtype=torch.bfloat16))
        if self._use_tma_params():
            buf['p_hi'] = torch.empty(M, n_par, device=device, dtype=torch.bfloat16)
            buf['p_lo'] = torch.empty(M, n_par, device=device, dtype=torch.bfloat16)
        else:
            buf['params'] = torch.empty(M, n_par, device=device, dtype=torch.float32)
        return buf

    def _use_tma_params(self):
        return self.tma_params and self.precision == 'bf16x3' and self.in_points == 32 and self.eff_in_dim == 64 and self.out_points == 128

    def generate_params(self, q2, buf, presplit=False):
        """Stage 1: dynamic mixing parameters [M, G*(C*C + Pout*Pin)] = query @ W^T + b on tcgen05 (depends on the query only,
        NOT on the sampled features -> can run concurrently with the gather)."""
        M, D = q2.shape
        x3 = self.precision == 'bf16x3'
        if not presplit:                  # (the decoder layer lets the producing dense chain write q_hi / q_lo directly)
            ops.split_bf16(q2, need_lo=x3, out=(buf['q_hi'], buf['q_lo'] if x3 else None))
        w_hi, w_lo = self._pg.get(self.parameter_generator.weight)
        if 'p_hi' in buf:
            ops.gemm_bf16_tn_split(buf['q_hi'], buf['q_lo'], w_hi, w_lo, M, self.n_groups * self.total_parameters, D,
                                   bias=self.parameter_generator.bias, out=(buf['p_hi'], buf['p_lo']))
            return buf['p_hi'], buf['p_lo']
        a, b = ([buf['q_hi'], buf['q_hi'], buf['q_lo']], [w_hi, w_lo, w_hi]) if x3 else ([buf['q_hi']], [w_hi])
        ops.gemm_bf16_tn(a, b, M, self.n_groups * self.total_parameters, D, bias=self.parameter_generator.bias, out=buf['params'])
        return buf['params']

    def mix_and_project(self, params, x, q2, norm=None, defer_reduce=False):
        """Stages 2+3: per-(query, group) mixing, out_proj (split-K tcgen05) and the fused reduce + residual + LayerNorm.
        defer_reduce: return the split-K partials and the reduce operands instead (dict for ops.dense_chain_reduce), so the
        caller's next dense chain performs the reduce + norm in its prologue."""
        M, G, P, C = x.shape
        D = self.query_dim
        if isinstance(params, tuple):
            y_hi, y_lo, _ = ops.mix_presplit(params[0], params[1], x)
        else:
            y_hi, y_lo, _ = ops.mix(params, x)
        o_hi, o_lo = self._op.get(self.out_proj.weight)
        K2 = self.out_proj.in_features
        a, b = ([y_hi, y_hi, y_lo], [o_hi, o_lo, o_hi]) if self.precision == 'bf16x3' else ([y_hi], [o_hi])
        split_k = max(1, min(self.split_k, K2 // 64))
        partial = ops.gemm_bf16_tn(a, b, M, D, K2, split_k=split_k)
        if defer_reduce:
            return dict(partial=partial if partial.dim() == 3 else partial[None], bias=self.out_proj.bias, residual=q2,
                        ln_w=None if norm is None else norm.weight, ln_b=None if norm is None else norm.bias)
        return ops.reduce_ln(partial, bias=self.out_proj.bias, residual=q2,
                             ln_w=None if norm is None else norm.weight, ln_b=None if norm is None else norm.bias)

    def forward_fused(self, x, query, norm=None):
        """x [B,Q,G,P,C], query [B,Q,D] -> norm(query + out_proj(mix(x; params(query)))) [B,Q,D]
        (`norm` = the LayerNorm applied right after in the decoder layer, fused into the split-K reduce)."""
        B, Q, G, P, C = x.shape
        assert G == self.n_groups and P == self.in_points and C == self.eff_in_dim
        M, D = B * Q, self.query_dim
        q2 = query.reshape(M, D)
        params = self.generate_params(q2, self.alloc_params(M, q2.device))
        return self.mix_and_project(params, x.reshape(M, G, P, C), q2, norm).reshape(B, Q, D)

    def forward(self, x, query):
        return self.forward_fused(x, query, None)


class _MHAParams(nn.Module):
    """Parameter holder named like mmcv's MultiheadAttention wrapper: `.attn` is an nn.MultiheadAttention."""

    def __init__(self, embed_dims, num_heads, dropout):
        super().__init__()
        self.attn = nn.MultiheadAttention(embed_dims, num_heads, dropout)


class SparseBEVSelfAttention(BaseModule):
    """Scale-adaptive Self Attention (reference :196-248)."""

    def __init__(self, embed_dims=256, num_heads=8, dropout=0.1, pc_range=[], init_cfg=None):
        super().__init__(init_cfg)
        self.pc_range = pc_range
        self.num_heads = num_heads
        self.attention = _MHAParams(embed_dims, num_heads, dropout)
        self.gen_tau = nn.Linear(embed_dims, num_heads)
        self._cache_in, self._cache_out = ops.DenseWeight(), ops.DenseWeight()
        self.core_impl = 'split'

    @torch.no_grad()
    def init_weights(self):
        nn.init.zeros_(self.gen_tau.weight)
        nn.init.uniform_(self.gen_tau.bias, 0.0, 2.0)

    def in_layer(self, y, y_hi=None, y_lo=None):
        """in_proj and gen_tau as ONE concatenated Linear: columns [0,3D) = q|k|v, [3D,3D+H) = tau; optionally also stored
        as the bf16 (hi, lo) pair the tensor-core attention core consumes."""
        attn = self.attention.attn
        wt, ldw, bias = self._cache_in.get_with_bias([attn.in_proj_weight, self.gen_tau.weight], [attn.in_proj_bias, self.gen_tau.bias])
        c = self._cache_in
        return ops.chain_layer(wt, ldw, attn.embed_dim, 3 * attn.embed_dim + self.num_heads, bias=bias, y=y, w_hi=c.w_hi, w_lo=c.w_lo, kpad=c.kpad,
                               y_hi=y_hi, y_lo=y_lo, w_pack=c.w_pack)

    def out_layer(self, residual, norm, y, y_hi=None, y_lo=None):
        attn = self.attention.attn
        wt, ldw, bias = self._cache_out.get_with_bias([attn.out_proj.weight], [attn.out_proj.bias])
        c = self._cache_out
        return ops.chain_layer(wt, ldw, attn.embed_dim, attn.embed_dim, bias=bias, ln=norm, residual=residual, res_pre_ln=True, y=y,
                               w_hi=c.w_hi, w_lo=c.w_lo, kpad=c.kpad, y_hi=y_hi, y_lo=y_lo, w_pack=c.w_pack)

    def attention_core(self, query_bbox, x, pre_attn_mask=None, pre=None):
        """x [B*Q, D] -> softmax(qk^T/sqrt(d) - tau*dist) v, heads concatenated [B*Q, D] (before out_proj).
        pre = (x0, ldx0, layers): dense layers chained IN FRONT of the in-projection in the same launch; the last of them
        must produce x (the decoder layer passes its position encoder here: one launch less)."""
        B, Q = query_bbox.shape[:2]
        D, H = self.attention.attn.embed_dim, self.num_heads
        qkvt = torch.empty(B * Q, 3 * D + H, device=x.device, dtype=torch.float32)
        x0, ldx0, head = (x, D, []) if pre is None else pre
        if self.core_impl == 'split':          # tensor-core path: the chain epilogue also emits the bf16 (hi, lo) operands
            hi = torch.empty(B * Q, 3 * D + H, device=x.device, dtype=torch.bfloat16)
            lo = torch.empty_like(hi)
            ops.dense_chain(x0, ldx0, B * Q, list(head) + [self.in_layer(qkvt, hi, lo)])
            if DUMP.enabled:               # reference :218-219
                torch.save(qkvt[:, 3 * D:3 * D + H].reshape(B, Q, H).cpu(), '{}/sasa_tau_stage{}.pth'.format(DUMP.out_dir, DUMP.stage_count))
            o = ops.sasa_split(qkvt, query_bbox, self.pc_range, H, D, dn_mask=pre_attn_mask, split=(hi, lo))
            return o.reshape(B * Q, D)
        ops.dense_chain(x0, ldx0, B * Q, list(head) + [self.in_layer(qkvt)])
        o = ops.sasa(qkvt, query_bbox, qkvt[:, 3 * D:], self.pc_range, H, dn_mask=pre_attn_mask,
                     ld_qkv=3 * D + H, ld_tau=3 * D + H, embed_dims=D)
        return o.reshape(B * Q, D)

    def forward_fused(self, query_bbox, query_feat, pre_attn_mask=None, norm=None):
        """-> norm(query_feat + out_proj(attention)) ; the [B*8,Q,Q] mask is never built."""
        B, Q, D = query_feat.shape
        x = query_feat.reshape(B * Q, D)
        o = self.attention_core(query_bbox, x, pre_attn_mask)
        out = torch.empty(B * Q, D, device=x.device, dtype=torch.float32)
        ops.dense_chain(o, D, B * Q, [self.out_layer(x, norm, out)])
        return out.reshape(B, Q, D)

    def forward(self, query_bbox, query_feat, pre_attn_mask):
        return self.forward_fused(query_bbox, query_feat, pre_attn_mask, None)


class SparseBEVSampling(BaseModule):
    """Adaptive Spatio-temporal Sampling (reference :251-317)."""

    def __init__(self, embed_dims=256, num_frames=4, num_groups=4, num_points=8, num_levels=4, pc_range=[], init_cfg=None):
        super().__init__(init_cfg)
        assert num_groups == ops.GROUPS
        self.num_frames, self.num_points, self.num_groups, self.num_levels = num_frames, num_points, num_groups, num_levels
        self.pc_range = pc_range
        self.sampling_offset = nn.Linear(embed_dims, num_groups * num_points * 3)
        self.scale_weights = nn.Linear(embed_dims, num_groups * num_points * num_levels)
        self._heads = _Dense(self.sampling_offset, extra=[self.scale_weights])      # one Linear: [offset | scale logits]
        self.feat_layout = 'grouped'

    def init_weights(self):
        bias = self.sampling_offset.bias.data.view(self.num_groups * self.num_points, 3)
        nn.init.zeros_(self.sampling_offset.weight)
        nn.init.uniform_(bias[:, 0:3], -0.5, 0.5)

    def heads_layer(self, y):
        return self._heads.layer(y=y)

    def points_args(self):
        """Arguments of ops.dense_chain_points for a chain that ends in heads_layer: the sample points and scale weights
        then come out of that launch's epilogue."""
        G, P, L = self.num_groups, self.num_points, self.num_levels
        return dict(pc_range=self.pc_range, GP=G * P, L=L, off_col=0, log_col=G * P * 3)

    def sample(self, query_bbox, heads_out, mlvl_feats, img_metas, frame_window=None, scatter_ptrs=None, points=None):
        """heads_out [B*Q, G*P*3 + G*P*L] = the concatenated sampling_offset | scale_weights Linear output, or
        points = (pts, scale_w) already produced by ops.dense_chain_points.
        frame_window / scatter_ptrs: frame-sharded forms, see ops.sampling4d_fused."""
        B, Q = query_bbox.shape[:2]
        image_h, image_w, _ = img_metas[0]['img_shape'][0]
        G, P, L = self.num_groups, self.num_points, self.num_levels
        if points is not None:
            pts, sw = points
        else:
            ld = heads_out.shape[1]
            pts, sw = ops.sample_points(query_bbox, heads_out, heads_out[:, G * P * 3:], self.pc_range, L,
                                        num_points_total=G * P, ld_off=ld, ld_log=ld)
        if DUMP.enabled:                   # reference models/sparsebev_sampling.py:82-86 (slow export path, a few torch ops)
            cam, valid = projected_sample_points(pts, query_bbox[..., 8:10], img_metas[0]['time_diff'], img_metas[0]['lidar2img'],
                                                 image_h, image_w, num_views=NUM_VIEWS)
            torch.save(cam.cpu(), '{}/sample_points_cam_stage{}.pth'.format(DUMP.out_dir, DUMP.stage_count))
            torch.save(valid.cpu(), '{}/sample_points_cam_valid_mask_stage{}.pth'.format(DUMP.out_dir, DUMP.stage_count))
        return ops.sampling4d_fused(mlvl_feats, pts, query_bbox, img_metas[0]['time_diff'], img_metas[0]['lidar2img'],
                                    sw.reshape(B, Q, G, P, L), image_h, image_w, num_frames=self.num_frames,
                                    num_views=NUM_VIEWS, layout=self.feat_layout,       # [B,Q,G,T*P,C]
                                    frame_window=frame_window, scatter_ptrs=scatter_ptrs)

    def forward(self, query_bbox, query_feat, mlvl_feats, img_metas):
        B, Q, D = query_feat.shape
        heads = torch.empty(B * Q, self._heads.out_features, device=query_feat.device, dtype=torch.float32)
        ops.dense_chain(query_feat.reshape(B * Q, D), D, B * Q, [self.heads_layer(heads)])
        return self.sample(query_bbox, heads, mlvl_feats, img_metas)


class _FFNParams(nn.Module):
    """Parameter holder keyed like mmcv's FFN: layers.0.0 = Linear(D, hidden), layers.1 = Linear(hidden, D)."""

    def __init__(self, embed_dims, feedforward_channels, ffn_drop):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(ffn_drop)),
            nn.Linear(feedforward_channels, embed_dims), nn.Dropout(ffn_drop))


class SparseBEVTransformerDecoderLayer(BaseModule):
    def __init__(self, embed_dims, num_frames=8, num_points=4, num_levels=4, num_classes=10, code_size=10,
                 num_cls_fcs=2, num_reg_fcs=2, pc_range=[], init_cfg=None):
        super().__init__(init_cfg)
        self.embed_dims, self.num_classes, self.code_size, self.pc_range = embed_dims, num_classes, code_size, pc_range
        self.position_encoder = nn.Sequential(
            nn.Linear(3, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True),
            nn.Linear(embed_dims, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True))
        self.self_attn = SparseBEVSelfAttention(embed_dims, num_heads=8, dropout=0.1, pc_range=pc_range)
        self.sampling = SparseBEVSampling(embed_dims, num_frames=num_frames, num_groups=4, num_points=num_points,
                                          num_levels=num_levels, pc_range=pc_range)
        self.mixing = AdaptiveMixing(in_dim=embed_dims, in_points=num_points * num_frames, n_groups=4, out_points=128)
        self.ffn = _FFNParams(embed_dims, 512, 0.1)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(embed_dims), nn.LayerNorm(embed_dims), nn.LayerNorm(embed_dims)
        cls_branch = []
        for _ in range(num_cls_fcs):
            cls_branch += [nn.Linear(embed_dims, embed_dims), nn.LayerNorm(embed_dims), nn.ReLU(inplace=True)]
        cls_branch.append(nn.Linear(embed_dims, num_classes))
        self.cls_branch = nn.Sequential(*cls_branch)
        reg_branch = []
        for _ in range(num_reg_fcs):
            reg_branch += [nn.Linear(embed_dims, embed_dims), nn.ReLU(inplace=True)]
        reg_branch.append(nn.Linear(embed_dims, code_size))
        self.reg_branch = nn.Sequential(*reg_branch)
        pe, cb, rb = self.position_encoder, self.cls_branch, self.reg_branch
        self._pe0, self._pe1 = _Dense(pe[0], pe[1]), _Dense(pe[3], pe[4])
        self._ffn0, self._ffn1 = _Dense(self.ffn.layers[0][0]), _Dense(self.ffn.layers[1], self.norm3)
        self._cls = [_Dense(cb[3 * i], cb[3 * i + 1]) for i in range(num_cls_fcs)] + [_Dense(cb[3 * num_cls_fcs])]
        self._reg = [_Dense(rb[2 * i]) for i in range(num_reg_fcs)] + [_Dense(rb[2 * num_reg_fcs])]
        self.overlap = True          # run independent kernels of a layer on a second stream (parallel graph branches)
        self.frame_shard = None      # dist.FrameShard: this rank holds / samples only its window of the T frames
        self.query_shard = None      # dist.QueryShard: ONE scene across ranks -- frames local, every query-side stage sharded over queries
        # split-K slices of the out-projection in query-sharded mode (None: chosen from the local row count; SBEV_QSHARD_SPLIT_K: A/B runs)
        self.qshard_split_k = int(os.environ['SBEV_QSHARD_SPLIT_K']) if os.environ.get('SBEV_QSHARD_SPLIT_K') else None
        # cls || reg as 16-row CTAs (SBEV_DENSE_WIDE_CTA: 57 + 57 CTAs, both chains resident at once) when they run on two streams.
        # OFF: measured on the B200 a 16-row CTA takes 34 us per chain against 15 us for 8 rows (the chain is bound by each warp's
        # instruction stream, which doubles), so the pair finishes later (317 vs 311 us per step); SBEV_WIDE_CTA_HEADS=1 turns it on
        self.wide_cta_heads = os.environ.get('SBEV_WIDE_CTA_HEADS', '0') == '1'
        # order of the gather and the parameter GEMM (see _forward_impl): 0 = side by side on two streams, 1 = one stream, gather first,
        # 2 = two streams, GEMM launched second; -1 (default) = 1 for the 8-frame configuration, 0 otherwise.  Measured on the B200 at
        # r50-T8 (tools/shots/r2_order.sh): the two kernels do not overlap -- each runs near the L2 throughput cap -- and the step is
        # 0.3049 ms with order 1 against 0.3077 ms with order 0; with one frame the gather is 7 us and hides under the GEMM
        self.phase_order = int(os.environ.get('SBEV_PHASE_ORDER', '-1'))
        self.use_cuda_graph = False  # replay the layer's launches as ONE CUDA graph (captured on first use per input signature)
        self._streams = {}
        self._graphs, self._graph_pool = collections.OrderedDict(), None
        self.register_load_state_dict_post_hook(lambda module, incompatible_keys: module.invalidate_caches())

    def invalidate_caches(self):
        """Drop every derived device copy of the parameters (transposed / bf16 (hi, lo) weights, concatenated biases) and every
        captured CUDA graph.  The caches are keyed on (data_ptr, _version), which covers optimizer steps, load_state_dict and
        .to(); it does NOT see in-place writes through `.data` (`p.data.copy_()`, EMA hooks, `bias.data.view(...)` in an init
        routine) -- those do not bump `_version`.  Called from init_weights, after load_state_dict and after _apply (.to /
        .cuda / .half ...); call it yourself after any manual `.data` surgery."""
        caches = [d.cache for d in (self._pe0, self._pe1, self._ffn0, self._ffn1, self.sampling._heads, *self._cls, *self._reg)]
        caches += [self.self_attn._cache_in, self.self_attn._cache_out, self.mixing._pg, self.mixing._op]
        for c in caches:
            c.key = None
        self.reset_graphs()

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        if hasattr(self, '_graphs'):
            self.invalidate_caches()
        return out

    def _side_stream(self, device):
        key = str(device)
        if self._streams.get(key) is None:
            self._streams[key] = torch.cuda.Stream(device=device)
        return self._streams[key]

    @torch.no_grad()
    def init_weights(self):
        self.self_attn.init_weights()
        self.sampling.init_weights()
        self.mixing.init_weights()
        nn.init.constant_(self.cls_branch[-1].bias, float(-np.log((1 - 0.01) / 0.01)))   # mmcv bias_init_with_prob(0.01)
        self.invalidate_caches()           # (SparseBEVSampling.init_weights writes through bias.data: no version bump)

    def refine_bbox(self, bbox_proposal, bbox_delta, time_diff):
        return ops.refine_bbox(bbox_proposal, bbox_delta, time_diff)

    @torch.no_grad()
    def _sample(self, query_bbox, heads, mlvl_feats, img_metas, points=None):
        """Sampled features [B,Q,G,T*P,C] of ALL frames.  Unsharded: one fused gather.  Frame-sharded
        (self.frame_shard = dist.FrameShard): mlvl_feats hold this rank's frames only; the rows of the other frames
        arrive from the peers (stored by their gather kernels over NVLink, or by one all-gather)."""
        shard = self.frame_shard
        if shard is None or shard.world == 1:
            return self.sampling.sample(query_bbox, heads, mlvl_feats, img_metas, points=points)
        if shard.exchange == 'p2p':
            B, Q = query_bbox.shape[:2]
            s = self.sampling
            buf, ptrs = shard.peer_buffers((B, Q, s.num_groups, s.num_frames * s.num_points, self.embed_dims // s.num_groups),
                                           query_bbox.device)
            s.sample(query_bbox, heads, mlvl_feats, img_metas, frame_window=shard.window, scatter_ptrs=ptrs, points=points)
            shard.peer_barrier()
            return buf
        local = self.sampling.sample(query_bbox, heads, mlvl_feats, img_metas, frame_window=shard.window, points=points)
        return shard.all_gather(local)

    MAX_GRAPHS = 4

    def reset_graphs(self):
        self._graphs.clear()

    @torch.no_grad()
    def _forward_graphed(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas, impl=None):
        """`use_cuda_graph`: the layer's launches (two streams, programmatic dependent launches included) are captured
        once per input signature and replayed.  The signature is everything the capture bakes in: shapes, the ADDRESSES
        of the feature maps (a steady-state inference loop gets the same blocks back from the caching allocator every
        frame; a new address simply captures another graph, the cache keeps the MAX_GRAPHS most recent), image size,
        parameter addresses + versions and the kernel-variant options.  Per call only the small tensors move: query_bbox, query_feat,
        time_diff, lidar2img (and the mask) are copied into the graph's static inputs -- host (pinned) or device sources
        alike -- and the three results are returned as copies of the static outputs."""
        meta = img_metas[0]
        impl = impl if impl is not None else self._forward_impl
        key = (impl.__name__, tuple(query_bbox.shape), tuple(query_feat.shape), tuple((f.data_ptr(), tuple(f.shape)) for f in mlvl_feats),
               self.sampling.feat_layout, None if attn_mask is None else tuple(attn_mask.shape),
               tuple(meta['img_shape'][0]), tuple(meta['time_diff'].shape), tuple(meta['lidar2img'].shape),
               hash(tuple((p.data_ptr(), p._version) for p in self.parameters())), _lib.options_epoch, self.overlap,
               self.mixing.precision, self.mixing.tma_params, self.mixing.split_k, self.self_attn.core_impl)
        entry = self._graphs.get(key)
        if entry is None:
            dev = mlvl_feats[0].device
            static = dict(qb=torch.empty(query_bbox.shape, device=dev), qf=torch.empty(query_feat.shape, device=dev),
                          td=torch.empty(meta['time_diff'].shape, device=dev), l2i=torch.empty(meta['lidar2img'].shape, device=dev),
                          mask=None if attn_mask is None else torch.empty(attn_mask.shape, device=dev, dtype=attn_mask.dtype))
            for name, src in (('qb', query_bbox), ('qf', query_feat), ('td', meta['time_diff']), ('l2i', meta['lidar2img']), ('mask', attn_mask)):
                if src is not None:
                    static[name].copy_(src, non_blocking=True)
            smeta = [dict(meta)]
            smeta[0]['time_diff'], smeta[0]['lidar2img'] = static['td'], static['l2i']
            # (no reference to the feature tensors is kept: holding them would stop the allocator from handing the same
            # blocks to the next frame, and a graph whose addresses are never passed again is simply never replayed)
            run = lambda: impl(static['qb'], static['qf'], mlvl_feats, static['mask'], smeta)      # noqa: E731
            warm = torch.cuda.Stream(device=dev)           # one eager pass off the capture: weight caches, function attributes
            warm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(warm):
                run()
            torch.cuda.current_stream().wait_stream(warm)
            if self._graph_pool is None:
                self._graph_pool = torch.cuda.graph_pool_handle()      # one pool for all graphs of this layer: they never run concurrently
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, pool=self._graph_pool):
                outs = run()
            entry = (graph, static, outs)
            self._graphs[key] = entry
            while len(self._graphs) > self.MAX_GRAPHS:
                self._graphs.popitem(last=False)
        else:
            self._graphs.move_to_end(key)
            static = entry[1]
            for name, src in (('qb', query_bbox), ('qf', query_feat), ('td', meta['time_diff']), ('l2i', meta['lidar2img']), ('mask', attn_mask)):
                if src is not None:
                    static[name].copy_(src, non_blocking=True)
        entry[0].replay()
        return tuple(o.clone() for o in entry[2])

    @torch.no_grad()
    def _forward_qshard(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas):
        """Query- and frame-sharded layer (dist.QueryShard; B must be 1): same inputs and outputs as _forward_impl on
        EVERY rank (full query_bbox / query_feat in, full query_feat / cls / bbox out), `mlvl_feats` holding this rank's
        frames only.  Stages are the same kernels as the unsharded layer run on the rank's own rows [q0, q1); rows of the
        results are bit-identical to the unsharded layer when the out-projection uses the same split-K count."""
        sh = self.query_shard
        B, Q, D = query_feat.shape
        if B != 1:
            raise RuntimeError('query-sharded decoder layer: batch must be 1 (got %d)' % B)
        dev = query_feat.device
        probe = getattr(self, '_probe', None)          # bench.py --breakdown: [(stage, event)] recorded on the main stream

        def mark(name):
            if probe is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                probe.append((name, ev))
        mark('start')
        smp, mixing = self.sampling, self.mixing
        G, P, L, T = smp.num_groups, smp.num_points, smp.num_levels, smp.num_frames
        GP, C = G * P, D // G
        qpr, q0, q1 = sh.partition(Q)
        Ml = q1 - q0
        ar = sh.arena([('points', (Q, GP, 3)), ('scale_w', (Q, GP, L)), ('sampled', (qpr, G, T * P, C)),
                       ('feat0', (Q, D)), ('feat1', (Q, D)), ('cls0', (Q, self.num_classes)), ('cls1', (Q, self.num_classes)),
                       ('box0', (Q, self.code_size)), ('box1', (Q, self.code_size))], dev)
        v = ar['views']
        par = sh.flip()
        out_feat, out_cls, out_box = v['feat%d' % par], v['cls%d' % par], v['box%d' % par]
        query_bbox = query_bbox.contiguous()
        qb2 = query_bbox.reshape(Q, -1)
        qf = query_feat.reshape(Q, D).contiguous()
        new = lambda m, n: torch.empty(m, n, device=dev, dtype=torch.float32)      # noqa: E731
        sl = slice(q0, q1)
        # (1) position encoder + in-projection for ALL rows (every query's K / V is needed by every rank)
        q1_all = new(Q, D)
        pos_enc = (qb2, qb2.shape[-1], [self._pe0.layer(relu=True), self._pe1.layer(relu=True, residual=qf, y=q1_all)])
        attn = self.self_attn
        H = attn.num_heads
        qkvt = new(Q, 3 * D + H)
        hi = torch.empty(Q, 3 * D + H, device=dev, dtype=torch.bfloat16)
        lo = torch.empty_like(hi)
        ops.dense_chain(pos_enc[0], pos_enc[1], Q, list(pos_enc[2]) + [attn.in_layer(qkvt, hi, lo)])
        mark('pos_enc+in_proj (all rows)')
        # (2) attention core of the own queries, out-projection + norm1 + sampling heads, sample points -> exchange 1
        o = new(Q, D)
        pbuf = mixing.alloc_params(max(Ml, 1), dev)
        q2, heads = new(max(Ml, 1), D), new(max(Ml, 1), smp._heads.out_features)
        main = torch.cuda.current_stream()
        side = self._side_stream(dev) if self.overlap else None
        params = None
        if Ml > 0:
            ops.sasa_split(qkvt, query_bbox, attn.pc_range, H, D, dn_mask=attn_mask, split=(hi, lo), q_range=(q0, q1), out=o.view(1, Q, D))
            mark('attention core (own queries)')
            ops.dense_chain(o[sl], D, Ml, [attn.out_layer(q1_all[sl], self.norm1, q2, pbuf['q_hi'], pbuf['q_lo']), smp.heads_layer(heads)])
            mark('out_proj+norm1+heads')
            # (4a) the parameter GEMM needs only q2: it forks off HERE, so it already runs while the sample points are computed and
            # exchanged (two kernels + one NVLink round trip), not only next to the gather
            if side is not None:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    params = mixing.generate_params(q2, pbuf, presplit=True)
            ld = heads.shape[1]
            ops.sample_points(query_bbox[:, sl], heads, heads[:, GP * 3:], smp.pc_range, L, num_points_total=GP, ld_off=ld, ld_log=ld,
                              out=(v['points'][sl], v['scale_w'][sl]))
            mark('sample_points')
        sh.exchange(ar, [('points', q0, q1), ('scale_w', q0, q1)])
        mark('exchange 1 (points)')
        # (3) gather: own frames, all queries, rows stored to the owning rank
        image_h, image_w, _ = img_metas[0]['img_shape'][0]
        meta = img_metas[0]
        if Ml > 0 and side is None:
            params = mixing.generate_params(q2, pbuf, presplit=True)
        ops.sampling4d_fused(mlvl_feats, v['points'].view(1, Q, GP, 3), query_bbox, meta['time_diff'], meta['lidar2img'],
                             v['scale_w'].view(1, Q, G, P, L), image_h, image_w, num_frames=T, num_views=NUM_VIEWS,
                             layout=smp.feat_layout, frame_window=sh.window, owner_ptrs=sh.peer_ptrs(ar, 'sampled'), q_per_rank=qpr)
        mark('gather (own frames, rows to owners)')
        sh.exchange(ar, [])
        mark('barrier')
        if Ml > 0 and side is not None:
            main.wait_stream(side)
        # (4b) mixing, (5) FFN, cls / reg + refine: own queries, results into this rank's rows of the output buffers -> exchange 2
        if Ml > 0:
            # split-K of the out-projection by local row count (tests/perf/kernel_sweep.py on the B200: the GEMM wants ~148 CTAs
            # of work, the reduce in front of the FFN chain wants few partials): 18 / 36 slices reduced in the FFN chain's
            # prologue as in the unsharded layer, 72 / 96 slices by the one-CTA-per-row reduce kernel
            keep_split = mixing.split_k
            if self.qshard_split_k is not None:
                mixing.split_k = self.qshard_split_k
            else:
                mixing.split_k = 96 if Ml <= 128 else 72 if Ml <= 256 else 36 if Ml <= 512 else keep_split      # (in-place A/B on the B200: tools/shots/r2_emu8.sh)
            separate_reduce = mixing.split_k > 36
            try:
                red = mixing.mix_and_project(params, v['sampled'][:Ml], q2, self.norm2, defer_reduce=True)
            finally:
                mixing.split_k = keep_split
            mark('wait param GEMM + mix + out_proj GEMM')
            q3 = new(Ml, D)
            td = meta['time_diff']
            q4, cls_score, bbox_pred = out_feat[sl], out_cls[sl], out_box[sl]
            cls_chain = [l.layer(relu=True) for l in self._cls[:-1]] + [self._cls[-1].layer(y=cls_score)]
            reg_chain = [l.layer(relu=True) for l in self._reg[:-1]] + [self._reg[-1].layer(refine=True, y=bbox_pred)]
            ffn_chain = [self._ffn0.layer(relu=True), self._ffn1.layer(residual=q3, res_pre_ln=True, y=q4)]
            refine = dict(refine_proposal=qb2[sl], refine_time_diff=td, refine_Q=Ml, refine_T=td.shape[1])
            if separate_reduce:
                q3r = ops.reduce_ln(red['partial'], red['bias'], red['residual'], red['ln_w'], red['ln_b'])
                ffn_chain = [self._ffn0.layer(relu=True), self._ffn1.layer(residual=q3r, res_pre_ln=True, y=q4)]
            if side is not None:
                if separate_reduce:
                    ops.dense_chain(q3r, D, Ml, ffn_chain)
                else:
                    ops.dense_chain_reduce(red['partial'], red['bias'], red['residual'], red['ln_w'], red['ln_b'], q3, ffn_chain)
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    ops.dense_chain(q4, D, Ml, reg_chain, **refine)
                mark('reduce+norm2+FFN')
                ops.dense_chain(q4, D, Ml, cls_chain)
                main.wait_stream(side)
            else:
                if separate_reduce:
                    ops.dense_chain(q3r, D, Ml, ffn_chain + cls_chain)
                else:
                    ops.dense_chain_reduce(red['partial'], red['bias'], red['residual'], red['ln_w'], red['ln_b'], q3, ffn_chain + cls_chain)
                ops.dense_chain(q4, D, Ml, reg_chain, **refine)
            mark('cls || reg+refine')
        sh.exchange(ar, [('feat%d' % par, q0, q1), ('cls%d' % par, q0, q1), ('box%d' % par, q0, q1)])
        mark('exchange 2 (outputs)')
        # (the symmetric output buffers are reused two layers later: hand out copies of the small results)
        return out_feat.view(1, Q, D), out_cls.view(1, Q, self.num_classes).clone(), out_box.view(1, Q, self.code_size).clone()

    def forward(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas):
        if DUMP.enabled:                   # export path: plain launches of the unsharded layer (files are written between kernels)
            if (self.query_shard is not None and self.query_shard.world > 1) or (self.frame_shard is not None and self.frame_shard.world > 1):
                raise RuntimeError('DUMP export needs the unsharded decoder layer')
            return self._forward_impl(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas)
        if self.query_shard is not None and self.query_shard.world > 1:
            if self.use_cuda_graph and not torch.cuda.is_current_stream_capturing():
                # (every rank captures at the same call: the exchanges inside the graph are plain kernels of ours)
                return self._forward_graphed(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas, impl=self._forward_qshard)
            return self._forward_qshard(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas)
        if (self.use_cuda_graph and (self.frame_shard is None or self.frame_shard.world == 1)
                and not torch.cuda.is_current_stream_capturing()):
            return self._forward_graphed(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas)
        return self._forward_impl(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas)

    def _forward_impl(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas):
        """query_bbox [B,Q,10] (cx,cy,cz,w,h,d,sin,cos,vx,vy normalised), query_feat [B,Q,D]
        -> (query_feat, cls_score [B,Q,num_classes], bbox_pred [B,Q,10])  (reference :162-193).

        11 kernel launches: 5 dense chains (two of them also emit the bf16 (hi, lo) operands of the tensor-core kernels
        that follow; the FFN chain's prologue performs the mixing stage's split-K reduce + norm2), SASA core, sample_points
        (or, option dense_fuse_points, the epilogue of the out-projection chain), fused gather, 2 tcgen05 GEMMs, mix; the gather runs
        concurrently with the parameter GEMM, cls with reg."""
        B, Q, D = query_feat.shape
        M, dev = B * Q, query_feat.device
        query_bbox = query_bbox.contiguous()
        qf = query_feat.reshape(M, D).contiguous()
        new = lambda n: torch.empty(M, n, device=dev, dtype=torch.float32)      # noqa: E731
        # (1) position encoder, + query_feat                                     [Linear(3,D) LN ReLU Linear LN ReLU] + residual
        q1 = new(D)
        pos_enc = (query_bbox.reshape(M, -1), query_bbox.shape[-1], [self._pe0.layer(relu=True), self._pe1.layer(relu=True, residual=qf, y=q1)])
        # (2) scale-adaptive self-attention: in-projection (+ tau) chained behind the position encoder, attention core,
        #     then out_proj + identity + norm1 chained with the sampling heads
        o = self.self_attn.attention_core(query_bbox, q1, attn_mask, pre=pos_enc)
        q2, heads = new(D), new(self.sampling._heads.out_features)
        pbuf = self.mixing.alloc_params(M, dev)               # q2 also leaves the chain as the bf16 (hi, lo) operand of the param GEMM
        chain2 = [self.self_attn.out_layer(q1, self.norm1, q2, pbuf['q_hi'], pbuf['q_lo']), self.sampling.heads_layer(heads)]
        if _lib.get_option('dense_fuse_points'):              # sample points + scale weights from the same launch (off by default: measured slower)
            points = ops.dense_chain_points(o, D, M, chain2, query_bbox, **self.sampling.points_args())
        else:
            points = None
            ops.dense_chain(o, D, M, chain2)
        # (3) adaptive spatio-temporal sampling  ||  (4a) dynamic-parameter GEMM: independent of each other (the GEMM needs
        # only q2), complementary resources (gather: LSU / L2 latency, no shared memory; GEMM: tensor cores + TMA) -> two
        # streams, i.e. two parallel branches when the layer is captured into a CUDA graph.  Buffers are allocated before the
        # fork so the caching allocator never sees cross-stream frees.
        main = torch.cuda.current_stream()
        side = self._side_stream(dev) if self.overlap else None
        order = self.phase_order if self.phase_order >= 0 else (1 if self.sampling.num_frames >= 8 else 0)
        if order == 1:                   # experiment: one stream, gather FIRST, so that the mix starts on the GEMM's heels (its last-written
            sampled = self._sample(query_bbox, heads, mlvl_feats, img_metas, points)          # parameter groups are still in L2)
            params = self.mixing.generate_params(q2, pbuf, presplit=True)
        elif order == 2 and side is not None:      # experiment: two streams, the GEMM launched BEHIND the sample-points kernel's start
            side.wait_stream(main)
            with torch.cuda.stream(side):
                sampled = self._sample(query_bbox, heads, mlvl_feats, img_metas, points)
            params = self.mixing.generate_params(q2, pbuf, presplit=True)
            main.wait_stream(side)
        elif side is not None:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                params = self.mixing.generate_params(q2, pbuf, presplit=True)
            sampled = self._sample(query_bbox, heads, mlvl_feats, img_metas, points)
            main.wait_stream(side)
        else:
            params = self.mixing.generate_params(q2, pbuf, presplit=True)
            sampled = self._sample(query_bbox, heads, mlvl_feats, img_metas, points)
        # (4b) adaptive mixing (+ identity + norm2)
        G, P = self.mixing.n_groups, self.mixing.in_points
        red = self.mixing.mix_and_project(params, sampled.reshape(M, G, P, -1), q2, self.norm2, defer_reduce=True)
        # (5) FFN (+ identity + norm3), its prologue finishing the mixing stage (split-K reduce + out_proj bias + identity + norm2
        #     -> q3); then classification and regression branches side by side
        q3, q4, cls_score, bbox_pred = new(D), new(D), new(self.num_classes), new(self.code_size)

        def ffn(chain):
            ops.dense_chain_reduce(red['partial'], red['bias'], red['residual'], red['ln_w'], red['ln_b'], q3, chain)
        td = img_metas[0]['time_diff']
        # cls || reg run side by side on two streams; with 8-row CTAs (113 + 113 on 148 SMs, one per SM) the second chain mostly runs in
        # the first one's wake (kernel timeline: 17 + 30 us); the 16-row form that would make both resident is slower still (see __init__)
        wide = side is not None and self.wide_cta_heads
        cls_chain = [l.layer(relu=True, wide_cta=wide and i == 0) for i, l in enumerate(self._cls[:-1])] + [self._cls[-1].layer(y=cls_score)]
        reg_chain = [l.layer(relu=True, wide_cta=wide and i == 0) for i, l in enumerate(self._reg[:-1])] + [self._reg[-1].layer(refine=True, y=bbox_pred)]
        ffn_chain = [self._ffn0.layer(relu=True), self._ffn1.layer(residual=q3, res_pre_ln=True, y=q4)]
        if side is not None:
            ffn(ffn_chain)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                ops.dense_chain(q4, D, M, reg_chain, refine_proposal=query_bbox, refine_time_diff=td, refine_Q=Q, refine_T=td.shape[1])
            ops.dense_chain(q4, D, M, cls_chain)
            main.wait_stream(side)
        else:
            ffn(ffn_chain + cls_chain)
            ops.dense_chain(q4, D, M, reg_chain, refine_proposal=query_bbox, refine_time_diff=td, refine_Q=Q, refine_T=td.shape[1])
        if DUMP.enabled:                   # reference :185-191
            torch.save(decode_bbox(query_bbox, self.pc_range).cpu(), '{}/query_bbox_stage{}.pth'.format(DUMP.out_dir, DUMP.stage_count))
            torch.save(decode_bbox(bbox_pred.reshape(B, Q, -1), self.pc_range).cpu(), '{}/bbox_pred_stage{}.pth'.format(DUMP.out_dir, DUMP.stage_count))
            torch.save(torch.sigmoid(cls_score.reshape(B, Q, -1)).cpu(), '{}/cls_score_stage{}.pth'.format(DUMP.out_dir, DUMP.stage_count))
        return q4.reshape(B, Q, D), cls_score.reshape(B, Q, self.num_classes), bbox_pred.reshape(B, Q, self.code_size)


class SparseBEVTransformerDecoder(BaseModule):
    def __init__(self, embed_dims, num_frames=8, num_points=4, num_layers=6, num_levels=4, num_classes=10,
                 code_size=10, pc_range=[], init_cfg=None):
        super().__init__(init_cfg)
        self.num_layers, self.pc_range = num_layers, pc_range
        self.use_cuda_graph = False      # capture all layers of a forward into ONE CUDA graph (unsharded inference; see forward)
        # params are shared across all decoder layers
        self.decoder_layer = SparseBEVTransformerDecoderLayer(
            embed_dims, num_frames, num_points, num_levels, num_classes, code_size, pc_range=pc_range)

    @torch.no_grad()
    def init_weights(self):
        self.decoder_layer.init_weights()

    @staticmethod
    def prepare_metas(img_metas, batch, device):
        """Host metadata -> device tensors, stored into img_metas[0] as the reference does (:60-70)."""
        ts = np.array([m['img_timestamp'] for m in img_metas], dtype=np.float64).reshape(batch, -1, NUM_VIEWS)
        time_diff = np.mean(ts[:, :1, :] - ts, axis=-1).astype(np.float32)
        img_metas[0]['time_diff'] = torch.from_numpy(time_diff).to(device)
        lidar2img = np.asarray([m['lidar2img'] for m in img_metas]).astype(np.float32)
        img_metas[0]['lidar2img'] = torch.from_numpy(lidar2img).to(device).contiguous()

    def prepare_feats(self, mlvl_feats):
        """L x [B,T*N,G*C,H,W] -> the layout the gather reads; mutates the list in place like the reference (:73-85).

        If a level is already channels-last in memory it is used as-is (zero copy, 'nhwc' addressing);
        otherwise it is regrouped to the reference's [B*T*G,N,H,W,C] with one permute copy."""
        layouts = set()
        for lvl, feat in enumerate(mlvl_feats):
            B, TN, GC, H, W = feat.shape
            nhwc = feat.permute(0, 1, 3, 4, 2)
            if nhwc.is_contiguous():
                mlvl_feats[lvl] = nhwc
                layouts.add('nhwc')
            else:
                N, T, G, C = NUM_VIEWS, TN // NUM_VIEWS, 4, GC // 4
                f = feat.reshape(B, T, N, G, C, H, W).permute(0, 1, 3, 2, 5, 6, 4)
                mlvl_feats[lvl] = f.reshape(B * T * G, N, H, W, C).contiguous()
                layouts.add('grouped')
        if len(layouts) != 1:
            raise RuntimeError('all feature levels must share one memory format')
        self.decoder_layer.sampling.feat_layout = layouts.pop()
        return mlvl_feats

    def _layers(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas):
        cls_scores, bbox_preds = [], []
        for i in range(self.num_layers):
            DUMP.stage_count = i
            query_feat, cls_score, bbox_pred = self.decoder_layer(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas)
            query_bbox = bbox_pred.clone().detach()
            cls_scores.append(cls_score)
            bbox_preds.append(bbox_pred)
        return torch.stack(cls_scores), torch.stack(bbox_preds)

    def _layers_one_graph(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas):
        # body of the decoder-level graph: plain launches of all layers (no per-layer graphs inside the capture or its warm-up pass)
        layer = self.decoder_layer
        keep, layer.use_cuda_graph = layer.use_cuda_graph, False
        try:
            return self._layers(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas)
        finally:
            layer.use_cuda_graph = keep

    @torch.no_grad()
    def forward(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas):
        self.prepare_metas(img_metas, query_bbox.shape[0], query_bbox.device)
        self.prepare_feats(mlvl_feats)
        layer = self.decoder_layer
        sharded = (layer.query_shard is not None and layer.query_shard.world > 1) or (layer.frame_shard is not None and layer.frame_shard.world > 1)
        if self.use_cuda_graph and not sharded and not DUMP.enabled and not torch.cuda.is_current_stream_capturing():
            # ONE graph for the whole decoder (num_layers x 11 launches): the small inputs are copied in and the two stacked results
            # copied out once per forward instead of once per layer (same signature cache as the per-layer graphs)
            return layer._forward_graphed(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas, impl=self._layers_one_graph)
        return self._layers(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas)


class SparseBEVTransformer(BaseModule):
    def __init__(self, embed_dims, num_frames=8, num_points=4, num_layers=6, num_levels=4, num_classes=10,
                 code_size=10, pc_range=[], init_cfg=None):
        assert init_cfg is None, 'To prevent abnormal initialization behavior, init_cfg is not allowed to be set'
        super().__init__(init_cfg=init_cfg)
        self.embed_dims = embed_dims
        self.pc_range = pc_range
        self.decoder = SparseBEVTransformerDecoder(embed_dims, num_frames, num_points, num_layers, num_levels,
                                                   num_classes, code_size, pc_range=pc_range)

    @torch.no_grad()
    def init_weights(self):
        self.decoder.init_weights()

    def shard_frames(self, shard):
        """Frame-sharded operation (not in the reference, whose only multi-GPU mode is DDP): `shard` is a dist.FrameShard;
        afterwards `mlvl_feats` passed to forward hold only this rank's frames, img_metas still describe all of them."""
        self.decoder.decoder_layer.frame_shard = shard
        return self

    def shard_queries(self, shard):
        """Query- and frame-sharded operation (strong scaling of ONE scene; not in the reference): `shard` is a
        dist.QueryShard (None switches it off).  `mlvl_feats` passed to forward then hold only this rank's frames, every
        rank passes the same queries / img_metas and every rank gets the full result."""
        self.decoder.decoder_layer.query_shard = shard
        return self

    def forward(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas):
        cls_scores, bbox_preds = self.decoder(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas)
        return torch.nan_to_num(cls_scores), torch.nan_to_num(bbox_preds)


if TRANSFORMER is not None:                               # same registry name as the reference (:16)
    try:
        TRANSFORMER.register_module(module=SparseBEVTransformer, force=True)
    except Exception:                                     # pragma: no cover
        pass
