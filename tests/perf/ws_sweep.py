"""Device times of the decoder layer's five dense chains, streaming kernel (dense_ws 0) vs the weights-stationary cluster kernel
(dense_ws 1; row tiles automatic / 16 / 32), at the row counts of the 1/2/4/8-GPU query shards, and of the whole layer (CUDA graph)
with either.  Launch queue pre-filled, CUDA events.  Writes gpurun_out/ws_sweep.json.  Development tool."""
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
    o = torch.randn(Q, D, device=dev)
    q2, heads = new(Q, D), new(Q, smp._heads.out_features)
    pbuf = mixing.alloc_params(Q, dev)
    td = meta['time_diff']
    modes = [('stream', 0, 0), ('ws', 1, 0), ('ws16', 1, 16), ('ws32', 1, 32)]
    for M in (900, 450, 225, 113):
        part = torch.randn(18, M, D, device=dev) * 0.1
        q3, q4, cls, box = new(M, D), torch.randn(M, D, device=dev), new(M, 10), new(M, 10)
        q3.normal_()
        ffn_chain = [layer._ffn0.layer(relu=True), layer._ffn1.layer(residual=q3, res_pre_ln=True, y=q4)]
        cls_chain = [l.layer(relu=True) for l in layer._cls[:-1]] + [layer._cls[-1].layer(y=cls)]
        reg_chain = [l.layer(relu=True) for l in layer._reg[:-1]] + [layer._reg[-1].layer(refine=True, y=box)]
        fns = {
            'A_posenc_inproj': lambda: ops.dense_chain(qb2, 10, M, [layer._pe0.layer(relu=True), layer._pe1.layer(relu=True, residual=x0, y=q1_all), attn.in_layer(qkvt, hi, lo)]),
            'B_outproj_heads': lambda: ops.dense_chain(o, D, M, [attn.out_layer(q1_all, layer.norm1, q2, pbuf['q_hi'], pbuf['q_lo']), smp.heads_layer(heads)]),
            'C_reduce18+ffn': lambda: ops.dense_chain_reduce(part, mixing.out_proj.bias, q2[:M], layer.norm2.weight, layer.norm2.bias, q3, ffn_chain),
            'C_ffn': lambda: ops.dense_chain(q3, D, M, ffn_chain),
            'D_cls': lambda: ops.dense_chain(q4, D, M, cls_chain),
            'E_reg': lambda: ops.dense_chain(q4, D, M, reg_chain, refine_proposal=qb2, refine_time_diff=td, refine_Q=M, refine_T=T),
        }
        for cname, fn in fns.items():
            for mname, ws, rt in modes:
                _lib.set_option('dense_ws', ws)
                _lib.set_option('dense_ws_rt', rt)
                rec('%s M%d %s' % (cname, M, mname), fn)
    _lib.set_option('dense_ws_rt', 0)
    # whole layer as one CUDA graph
    layer.use_cuda_graph = True
    for ws in (0, 1, 2):
        _lib.set_option('dense_ws', ws)
        layer.reset_graphs()
        rec('layer graph dense_ws=%d' % ws, lambda: layer(qb, qf, feats, None, metas), iters=50, warm=5)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'ws_sweep.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
