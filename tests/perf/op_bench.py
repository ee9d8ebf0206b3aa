"""Op-level timing on one GPU: our msmv_sampling (op boundary + fused front-end) vs the UNMODIFIED reference CUDA
kernel compiled for sm_100a (oracle/_ref), same inputs.  Writes gpurun_out/op_bench.json.  Test infrastructure."""
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sparsebev_b200 import ops, synthetic as S  # noqa: E402


def ref_mod():
    so = os.path.join(ROOT, 'oracle', '_ref', '_msmv_sampling_cuda.so')
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location('_msmv_sampling_cuda', so)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def timeit(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ref = ref_mod()
    dev = torch.device('cuda:0')
    res = {}
    for name, T, dist in [('r50_704x256', 8, 'uniform'), ('r50_704x256', 8, 'realistic'), ('r50_704x256', 1, 'uniform'),
                          ('r101_1408x512', 8, 'uniform')]:
        cfg = S.layer_cfg(name, T)
        L, Q, P, G = cfg['num_levels'], cfg['num_query'], 4, 4
        Bp = T * G
        torch.manual_seed(0)
        feats = [torch.randn(Bp, 6, h, w, 64, device=dev) for h, w in cfg['levels']]
        if dist == 'uniform':
            loc = torch.rand(Bp, Q, P, 3, device=dev)
            loc[..., 2] = torch.randint(0, 6, (Bp, Q, P), device=dev).float() / 5
        else:   # realistic: project head-initialised boxes through the synthetic rig with the fused kernel, reuse its loc
            l2i, stamps = S.camera_rig(T, cfg['image_h'], cfg['image_w'])
            qb = S.init_query_bbox(Q, seed=2)[None].to(dev)
            off = (torch.rand(1, Q, G * P * 3, device=dev) - 0.5)
            pts, sw = ops.sample_points(qb.contiguous(), off, torch.randn(1, Q, G * P * L, device=dev), cfg['pc_range'], L)
            td = torch.arange(T, device=dev, dtype=torch.float32)[None] * 0.5
            _, loc = ops.sampling4d_fused(feats, pts, qb[..., 8:10].contiguous(), td, l2i[None].contiguous().to(dev),
                                          sw.reshape(1, Q, G, P, L), cfg['image_h'], cfg['image_w'], num_frames=T, return_loc=True)
        w = torch.softmax(torch.randn(Bp, Q, P, L, device=dev), -1)
        key = '%s_T%d_%s' % (name, T, dist)
        ours = timeit(lambda: ops.msmv_forward(feats, loc, w))
        entry = {'ours_op_ms': ours, 'points': Bp * Q * P,
                 'algorithmic_MB': Bp * Q * P * (L * 4 * 64 * 4 + (3 + L) * 4 + 64 * 4) / 1e6}
        entry['ours_GBps'] = entry['algorithmic_MB'] / ours
        if ref is not None:
            f = ref._ms_deform_attn_cuda_c2345_forward if L == 4 else ref._ms_deform_attn_cuda_c23456_forward
            t = timeit(lambda: f(*feats, loc, w))
            entry.update(ref_cuda_ms=t, speedup_vs_ref_cuda=t / ours)
            go = torch.randn(Bp, Q, 64, P, device=dev)
            fb = ref._ms_deform_attn_cuda_c2345_backward if L == 4 else ref._ms_deform_attn_cuda_c23456_backward
            entry['ref_cuda_bwd_ms'] = timeit(lambda: fb(go, *feats, loc, w), iters=10, warm=2)
            entry['ours_bwd_ms'] = timeit(lambda: ops.msmv_backward(go, feats, loc, w), iters=10, warm=2)
            entry['ours_bwd_deterministic_ms'] = timeit(lambda: ops.msmv_backward(go, feats, loc, w, deterministic=True), iters=10, warm=2)
        res[key] = entry
        print(key, json.dumps(entry), flush=True)
        del feats
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'op_bench.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
