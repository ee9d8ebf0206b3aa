"""Phase timeline of the weights-stationary chain kernel (sbev_dense_chain_ws_debug: clock64 stamps of thread 0 of every CTA) for
the decoder layer's chains at a given row count.  usage: ws_timeline.py [M ...]   Development tool."""
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
    lib = _lib.load()
    _lib.set_option('dense_ws', 1)
    stamps = torch.zeros(256 * 64, dtype=torch.int64, device=dev)
    for M in [int(a) for a in sys.argv[1:] if a.isdigit()] or [113, 900]:
        part = torch.randn(18, M, D, device=dev) * 0.1
        q3, q4, cls, box = torch.randn(M, D, device=dev), torch.randn(M, D, device=dev), new(M, 10), new(M, 10)
        ffn_chain = [layer._ffn0.layer(relu=True), layer._ffn1.layer(residual=q3, res_pre_ln=True, y=q4)]
        cls_chain = [l.layer(relu=True) for l in layer._cls[:-1]] + [layer._cls[-1].layer(y=cls)]
        fns = {
            'A_posenc_inproj': lambda: ops.dense_chain(qb2, 10, M, [layer._pe0.layer(relu=True), layer._pe1.layer(relu=True, residual=x0, y=q1_all), attn.in_layer(qkvt, hi, lo)]),
            'C_reduce18+ffn': lambda: ops.dense_chain_reduce(part, mixing.out_proj.bias, q2[:M], layer.norm2.weight, layer.norm2.bias, q3, ffn_chain),
            'D_cls': lambda: ops.dense_chain(q4, D, M, cls_chain),
        }
        for cname, fn in fns.items():
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            stamps.zero_()
            lib.sbev_dense_chain_ws_debug(stamps.data_ptr())
            fn()
            torch.cuda.synchronize()
            lib.sbev_dense_chain_ws_debug(None)
            st = stamps.view(256, 64).cpu()
            print('==', cname, 'M', M)
            for cta in (0, 1, 7, 8, 15 * 8):
                row = [int(v) for v in st[cta] if int(v) != 0]
                if len(row) > 1:
                    print('cta %3d  total %6d cyc  deltas %s' % (cta, row[-1] - row[0], ' '.join(str(b - a) for a, b in zip(row[:-1], row[1:]))))


if __name__ == '__main__':
    main()
