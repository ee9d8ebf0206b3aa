"""Co-residency experiment: how long does the fused gather take while a persistent 1-CTA-per-SM kernel that holds S bytes of shared
memory (tests/perf/csrc/hog.cu, stands in for the parameter GEMM: 192 threads, 230 400 B today) occupies every SM?
gather alone ~44 us; ~(hog time + 44) means NO gather CTA fits next to the hog.  Development tool."""
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
    import ctypes
    so = os.path.join(ROOT, 'tests', 'perf', 'csrc', 'libhog.so')
    if not os.path.exists(so):                 # (built in-tree so that it travels with the gpurun snapshot; git-ignored)
        import subprocess
        subprocess.check_call([os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc'), '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-shared',
                               '-Xcompiler', '-fPIC', '-o', so, os.path.join(ROOT, 'tests', 'perf', 'csrc', 'hog.cu'), '-lcudart'])
    hog = ctypes.CDLL(so)
    hog.hog_launch.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p]
    smp = layer.sampling
    heads = torch.randn(Q, smp._heads.out_features, device=dev) * 0.1
    pts, sw = ops.sample_points(qb, heads, heads[:, 48:], cfg['pc_range'], L, num_points_total=16, ld_off=112, ld_log=112)
    sw5 = sw.reshape(1, Q, G, P, L)
    out_buf = torch.empty(1, Q, G, T * P, 64, device=dev)

    def gather():
        ops.sampling4d_fused(feats, pts, qb, meta['time_diff'], meta['lidar2img'], sw5, 256, 704, num_frames=T, layout='nhwc', out=out_buf)
    for _ in range(3):
        gather()
    torch.cuda.synchronize()
    sa, sb_ = torch.cuda.Stream(), torch.cuda.Stream()
    hog_us = 300.0
    cycles = int(hog_us * 1.9e3)
    for smem in (0, 230400, 229376, 228352, 227328, 226304, 224256, 221184, 217088, 200704, 163840, 98304):
        times = []
        for rep in range(5):
            torch.cuda.synchronize()
            e0, e1, h0, h1 = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if smem:
                with torch.cuda.stream(sa):
                    h0.record()
                    rc = hog.hog_launch(148, smem, cycles, ctypes.c_void_p(sa.cuda_stream))
                    h1.record()
                    assert rc == 0, rc
                torch.cuda._sleep(20000)                 # (the hog is resident before the gather is launched)
                sb_.wait_event(h0)
            with torch.cuda.stream(sb_):
                torch.cuda._sleep(40000)
                e0.record()
                gather()
                e1.record()
            torch.cuda.synchronize()
            times.append(1e3 * e0.elapsed_time(e1))
        print('hog smem %6d B (+1 KB reserved +static): gather %s us  (hog %.0f us)' % (smem, ' '.join('%.1f' % t for t in times), 1e3 * h0.elapsed_time(h1) if smem else 0), flush=True)


if __name__ == '__main__':
    main()
