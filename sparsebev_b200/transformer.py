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
hand-written sm_100a kernels in libsparsebev_b200.so through `ops` -- about two dozen launches per layer,
capturable in one CUDA graph (see `SparseBEVTransformerDecoder.forward_graphed`).  Forward only: this
is the eval path (dropout = identity, no activation checkpointing); training the decoder through these
modules is out of scope (the op-level autograd Functions in `wrapper.py` do have a backward).
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops

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


class _Dense:
    """Linear (+LN +ReLU +residual) launcher bound to the parameters of one nn.Linear."""

    def __init__(self, linear, ln=None):
        self.linear, self.ln, self.cache = linear, ln, ops.DenseWeight()

    def __call__(self, x, relu=False, residual=None, res_pre_ln=False, k=None):
        wt, ldw = self.cache.get(self.linear.weight)
        return ops.dense(x, wt, ldw, self.linear.out_features, bias=self.linear.bias,
                         ln_w=None if self.ln is None else self.ln.weight, ln_b=None if self.ln is None else self.ln.bias,
                         residual=residual, relu=relu, res_pre_ln=res_pre_ln, k=k)


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
        self.split_k = 16
        self._pg, self._op = _SplitWeight(), _SplitWeight()

    @torch.no_grad()
    def init_weights(self):
        nn.init.zeros_(self.parameter_generator.weight)

    def forward_fused(self, x, query, norm=None):
        """x [B,Q,G,P,C], query [B,Q,D] -> norm(query + out_proj(mix(x; params(query)))) [B,Q,D]
        (`norm` = the LayerNorm applied right after in the decoder layer, fused into the split-K reduce)."""
        B, Q, G, P, C = x.shape
        assert G == self.n_groups and P == self.in_points and C == self.eff_in_dim
        M, D = B * Q, self.query_dim
        q2 = query.reshape(M, D)
        q_hi, q_lo = ops.split_bf16(q2, need_lo=self.precision == 'bf16x3')
        w_hi, w_lo = self._pg.get(self.parameter_generator.weight)
        n_par = self.n_groups * self.total_parameters
        if self.precision == 'bf16x3':
            a, b = [q_hi, q_hi, q_lo], [w_hi, w_lo, w_hi]
        else:
            a, b = [q_hi], [w_hi]
        params = ops.gemm_bf16_tn(a, b, M, n_par, D, bias=self.parameter_generator.bias)
        y_hi, y_lo, _ = ops.mix(params, x.reshape(M, G, P, C))
        o_hi, o_lo = self._op.get(self.out_proj.weight)
        K2 = self.out_proj.in_features
        if self.precision == 'bf16x3':
            a, b = [y_hi, y_hi, y_lo], [o_hi, o_lo, o_hi]
        else:
            a, b = [y_hi], [o_hi]
        split_k = self.split_k
        while (K2 // 64) % split_k:
            split_k //= 2
        partial = ops.gemm_bf16_tn(a, b, M, D, K2, split_k=split_k)
        out = ops.reduce_ln(partial, bias=self.out_proj.bias, residual=q2,
                            ln_w=None if norm is None else norm.weight, ln_b=None if norm is None else norm.bias)
        return out.reshape(B, Q, D)

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
        self._in, self._out, self._tau = ops.DenseWeight(), ops.DenseWeight(), _Dense(self.gen_tau)

    @torch.no_grad()
    def init_weights(self):
        nn.init.zeros_(self.gen_tau.weight)
        nn.init.uniform_(self.gen_tau.bias, 0.0, 2.0)

    def forward_fused(self, query_bbox, query_feat, pre_attn_mask=None, norm=None):
        """-> norm(query_feat + out_proj(attention)) ; the [B*8,Q,Q] mask is never built."""
        B, Q, D = query_feat.shape
        x = query_feat.reshape(B * Q, D)
        attn = self.attention.attn
        wt, ldw = self._in.get(attn.in_proj_weight)
        qkv = ops.dense(x, wt, ldw, 3 * D, bias=attn.in_proj_bias)
        tau = self._tau(x)
        o = ops.sasa(qkv.reshape(B, Q, 3 * D), query_bbox, tau.reshape(B, Q, self.num_heads), self.pc_range,
                     self.num_heads, dn_mask=pre_attn_mask)
        wt, ldw = self._out.get(attn.out_proj.weight)
        out = ops.dense(o.reshape(B * Q, D), wt, ldw, D, bias=attn.out_proj.bias, residual=x, res_pre_ln=True,
                        ln_w=None if norm is None else norm.weight, ln_b=None if norm is None else norm.bias)
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
        self._off, self._sw = _Dense(self.sampling_offset), _Dense(self.scale_weights)
        self.feat_layout = 'grouped'

    def init_weights(self):
        bias = self.sampling_offset.bias.data.view(self.num_groups * self.num_points, 3)
        nn.init.zeros_(self.sampling_offset.weight)
        nn.init.uniform_(bias[:, 0:3], -0.5, 0.5)

    def forward(self, query_bbox, query_feat, mlvl_feats, img_metas):
        B, Q, D = query_feat.shape
        image_h, image_w, _ = img_metas[0]['img_shape'][0]
        x = query_feat.reshape(B * Q, D)
        offset = self._off(x)                                   # [BQ, G*P*3]
        logits = self._sw(x)                                    # [BQ, G*P*L]
        G, P, L = self.num_groups, self.num_points, self.num_levels
        pts, sw = ops.sample_points(query_bbox, offset.reshape(B, Q, G * P * 3), logits.reshape(B, Q, G * P * L),
                                    self.pc_range, L)
        vel = query_bbox[..., 8:10].contiguous()
        return ops.sampling4d_fused(mlvl_feats, pts, vel, img_metas[0]['time_diff'], img_metas[0]['lidar2img'],
                                    sw.reshape(B, Q, G, P, L), image_h, image_w, num_frames=self.num_frames,
                                    num_views=NUM_VIEWS, layout=self.feat_layout)       # [B,Q,G,T*P,C]


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

    @torch.no_grad()
    def init_weights(self):
        self.self_attn.init_weights()
        self.sampling.init_weights()
        self.mixing.init_weights()
        nn.init.constant_(self.cls_branch[-1].bias, float(-np.log((1 - 0.01) / 0.01)))   # mmcv bias_init_with_prob(0.01)

    def refine_bbox(self, bbox_proposal, bbox_delta, time_diff):
        return ops.refine_bbox(bbox_proposal, bbox_delta, time_diff)

    @torch.no_grad()
    def forward(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas):
        """query_bbox [B,Q,10] (cx,cy,cz,w,h,d,sin,cos,vx,vy normalised), query_feat [B,Q,D]
        -> (query_feat, cls_score [B,Q,num_classes], bbox_pred [B,Q,10])  (reference :162-193)."""
        B, Q, D = query_feat.shape
        M = B * Q
        query_bbox = query_bbox.contiguous()
        qf = query_feat.reshape(M, D).contiguous()
        h = self._pe0(query_bbox.reshape(M, -1), relu=True, k=3)                  # Linear(3,D) on bbox[..., :3] + LN + ReLU
        qf = self._pe1(h, relu=True, residual=qf)                                  # + LN + ReLU, then query_feat + query_pos
        qf = self.self_attn.forward_fused(query_bbox, qf.reshape(B, Q, D), attn_mask, self.norm1)
        sampled = self.sampling(query_bbox, qf, mlvl_feats, img_metas)
        qf = self.mixing.forward_fused(sampled, qf, self.norm2).reshape(M, D)
        h = self._ffn0(qf, relu=True)
        qf = self._ffn1(h, residual=qf, res_pre_ln=True)                           # identity + ffn, then norm3
        c = qf
        for layer in self._cls[:-1]:
            c = layer(c, relu=True)
        cls_score = self._cls[-1](c).reshape(B, Q, self.num_classes)
        r = qf
        for layer in self._reg[:-1]:
            r = layer(r, relu=True)
        delta = self._reg[-1](r).reshape(B, Q, self.code_size)
        bbox_pred = self.refine_bbox(query_bbox, delta, img_metas[0]['time_diff'])
        return qf.reshape(B, Q, D), cls_score, bbox_pred


class SparseBEVTransformerDecoder(BaseModule):
    def __init__(self, embed_dims, num_frames=8, num_points=4, num_layers=6, num_levels=4, num_classes=10,
                 code_size=10, pc_range=[], init_cfg=None):
        super().__init__(init_cfg)
        self.num_layers, self.pc_range = num_layers, pc_range
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

    @torch.no_grad()
    def forward(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas):
        self.prepare_metas(img_metas, query_bbox.shape[0], query_bbox.device)
        self.prepare_feats(mlvl_feats)
        cls_scores, bbox_preds = [], []
        for _ in range(self.num_layers):
            query_feat, cls_score, bbox_pred = self.decoder_layer(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas)
            query_bbox = bbox_pred.clone().detach()
            cls_scores.append(cls_score)
            bbox_preds.append(bbox_pred)
        return torch.stack(cls_scores), torch.stack(bbox_preds)


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

    def forward(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas):
        cls_scores, bbox_preds = self.decoder(query_bbox, query_feat, mlvl_feats, attn_mask, img_metas)
        return torch.nan_to_num(cls_scores), torch.nan_to_num(bbox_preds)


if TRANSFORMER is not None:                               # same registry name as the reference (:16)
    try:
        TRANSFORMER.register_module(module=SparseBEVTransformer, force=True)
    except Exception:                                     # pragma: no cover
        pass
