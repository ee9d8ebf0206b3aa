"""Per-kernel device times of the decoder layer's stages at the row counts of the 1/2/4/8-GPU query shards, for every
kernel-variant option worth comparing (dense_nsplit, gather_variant, split-K of the out-projection).  Launch queue
pre-filled (torch.cuda._sleep), CUDA events: back-to-back device time, warm L2 for the small operands.
Writes gpurun_out/kernel_sweep.json.  Test infrastructure / development tool."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import sparsebev_b200 as sb                                    # noqa: E402
from sparsebev_b200 import _lib, ops, synthetic as S           # noqa: E402


def dev_ms(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    torch.cuda._sleep(int((0.06 * iters + 0.3) * 1.9e6))
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / iters        # us


def main():
    only = set(sys.argv[1:])
    dev = torch.device('cuda:0')
    T, name = 8, 'r50_704x256'
    cfg = S.layer_cfg(name, T, num_layers=1)
    model = sb.SparseBEVTransformer(256, num_frames=T, num_points=4, num_layers=1, num_levels=4, pc_range=cfg['pc_range'])
    model.load_state_dict({'decoder.decoder_layer.' + k: v for k, v in S.make_state_dict(cfg, seed=0).items()})
    model = model.to(dev).eval()
    layer = model.decoder.decoder_layer
    Q, D, G, P, L = 900, 256, 4, 4, 4
    feats = model.decoder.prepare_feats([f.to(dev) for f in S.make_feats(name, T, batch=1, seed=100, memory_format='nhwc')])
    metas = S.make_metas(name, T, batch=1)
    model.decoder.prepare_metas(metas, 1, dev)
    meta = metas[0]
    qb = S.init_query_bbox(Q, seed=2)[None].contiguous().to(dev)
    qf = torch.randn(1, Q, D, generator=torch.Generator().manual_seed(3)).to(dev)
    res = {}

    def rec(key, fn, **kw):
        if only and not any(o in key for o in only):
            return
        try:
            res[key] = round(dev_ms(fn, **kw), 2)
        except Exception as e:                     # noqa
            res[key] = 'ERR %r' % (e,)
        print(key, res[key], flush=True)

    new = lambda m, n, dt=torch.float32: torch.empty(m, n, device=dev, dtype=dt)      # noqa: E731
    qb2, x0 = qb.reshape(Q, 10), qf.reshape(Q, D)
    attn, smp, mixing = layer.self_attn, layer.sampling, layer.mixing
    q1_all, qkvt = new(Q, D), new(Q, 3 * D + 8)
    hi, lo = new(Q, 3 * D + 8, torch.bfloat16), new(Q, 3 * D + 8, torch.bfloat16)

    def chain_a(M):
        ops.dense_chain(qb2, 10, M, [layer._pe0.layer(relu=True), layer._pe1.layer(relu=True, residual=x0, y=q1_all), attn.in_layer(qkvt, hi, lo)])
    chain_a(Q)
    o = new(Q, D)
    for ns in (0, 2, 4):
        _lib.set_option('dense_nsplit', ns)
        for M in (900, 450, 225, 113):
            rec('chainA_posenc_inproj ns%d M%d' % (ns, M), lambda: chain_a(M))
    _lib.set_option('dense_nsplit', 0)
    for M in (900, 450, 225, 113):
        rec('sasa M%d' % M, lambda: ops.sasa_split(qkvt, qb, cfg['pc_range'], 8, D, split=(hi, lo), q_range=(0, M), out=o.view(1, Q, D)))
    q2, heads = new(Q, D), new(Q, smp._heads.out_features)
    pbuf = mixing.alloc_params(Q, dev)

    def chain_b(M):
        ops.dense_chain(o, D, M, [attn.out_layer(q1_all, layer.norm1, q2, pbuf['q_hi'], pbuf['q_lo']), smp.heads_layer(heads)])
    chain_b(Q)
    for ns in (0, 2, 4):
        _lib.set_option('dense_nsplit', ns)
        for M in (900, 450, 225, 113):
            rec('chainB_outproj_heads ns%d M%d' % (ns, M), lambda: chain_b(M))
    _lib.set_option('dense_nsplit', 0)
    pts, sw = ops.sample_points(qb, heads, heads[:, 48:], cfg['pc_range'], L, num_points_total=16, ld_off=112, ld_log=112)
    rec('sample_points M900', lambda: ops.sample_points(qb, heads, heads[:, 48:], cfg['pc_range'], L, num_points_total=16, ld_off=112, ld_log=112))
    sw5 = sw.reshape(1, Q, G, P, L)
    for var in (2, 4, 5, 6, 1):
        _lib.set_option('gather_variant', var)
        for Tl in (8, 4, 2, 1):
            out_buf = torch.empty(1, Q, G, Tl * P, 64, device=dev)
            fw = [f[:, :Tl * 6].contiguous() for f in feats] if Tl < T else feats
            rec('gather v%d frames%d' % (var, Tl), lambda: ops.sampling4d_fused(fw, pts, qb, meta['time_diff'], meta['lidar2img'], sw5, 256, 704, num_frames=T,
                                                                                layout='nhwc', out=out_buf, frame_window=(0, Tl)))
    _lib.set_option('gather_variant', 2)
    sampled = ops.sampling4d_fused(feats, pts, qb, meta['time_diff'], meta['lidar2img'], sw5, 256, 704, num_frames=T, layout='nhwc')
    x4 = sampled.reshape(Q, G, T * P, 64)
    for M in (900, 450, 225, 113):
        pb = mixing.alloc_params(M, dev)
        pb['q_hi'].copy_(pbuf['q_hi'][:M]); pb['q_lo'].copy_(pbuf['q_lo'][:M])
        rec('param_gemm M%d' % M, lambda: mixing.generate_params(q2[:M], pb, presplit=True))
        params = mixing.generate_params(q2[:M], pb, presplit=True)
        yh, yl = new(M, 32768, torch.bfloat16), new(M, 32768, torch.bfloat16)
        rec('mix M%d' % M, lambda: ops.mix_presplit(params[0], params[1], x4[:M], out=(yh, yl)))
        oh, ol = mixing._op.get(mixing.out_proj.weight)
        for sk in (18, 36, 72, 128):
            if M == 900 and sk > 36:
                continue
            part = torch.empty(sk, M, D, device=dev)
            for impl in (0, 3):
                _lib.set_option('gemm_impl', impl)
                rec('out_gemm M%d splitk%d impl%d' % (M, sk, impl), lambda: ops.gemm_bf16_tn([yh, yh, yl], [oh, ol, oh], M, D, 32768, split_k=sk, out=part))
            _lib.set_option('gemm_impl', 0)
            q3, q4, cls = new(M, D), new(M, D), new(M, 10)
            ffn_chain = [layer._ffn0.layer(relu=True), layer._ffn1.layer(residual=q3, res_pre_ln=True, y=q4)]
            cls_chain = [l.layer(relu=True) for l in layer._cls[:-1]] + [layer._cls[-1].layer(y=cls)]
            for ns in (0, 4):
                _lib.set_option('dense_nsplit', ns)
                rec('reduce+ffn M%d splitk%d ns%d' % (M, sk, ns), lambda: ops.dense_chain_reduce(part, mixing.out_proj.bias, q2[:M], layer.norm2.weight, layer.norm2.bias, q3, ffn_chain))
            _lib.set_option('dense_nsplit', 0)
            rec('reduce_ln(separate) M%d splitk%d' % (M, sk), lambda: ops.reduce_ln(part, mixing.out_proj.bias, q2[:M], layer.norm2.weight, layer.norm2.bias))
        q3, q4, cls, box = new(M, D), new(M, D), new(M, 10), new(M, 10)
        ffn_chain = [layer._ffn0.layer(relu=True), layer._ffn1.layer(residual=q3, res_pre_ln=True, y=q4)]
        cls_chain = [l.layer(relu=True) for l in layer._cls[:-1]] + [layer._cls[-1].layer(y=cls)]
        reg_chain = [l.layer(relu=True) for l in layer._reg[:-1]] + [layer._reg[-1].layer(refine=True, y=box)]
        td = meta['time_diff']
        for ns in (0, 2, 4):
            _lib.set_option('dense_nsplit', ns)
            rec('ffn M%d ns%d' % (M, ns), lambda: ops.dense_chain(q3, D, M, ffn_chain))
            rec('cls M%d ns%d' % (M, ns), lambda: ops.dense_chain(q4, D, M, cls_chain))
            rec('reg M%d ns%d' % (M, ns), lambda: ops.dense_chain(q4, D, M, reg_chain, refine_proposal=qb2, refine_time_diff=td, refine_Q=M, refine_T=T))
        _lib.set_option('dense_nsplit', 0)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'kernel_sweep.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
