"""Context baseline: the reference decoder layer as STOCK PYTORCH EAGER on the B200,
calling the UNMODIFIED reference CUDA op (oracle/_ref), fp32 (TF32 off, like the reference).  This is what the
reference itself would run per decoder layer on this GPU; the layer is the oracle's restatement (mmcv glue replaced by
its documented semantics) executed on CUDA tensors.  Test infrastructure; writes gpurun_out/ref_layer_gpu.json."""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_torch as R              # noqa: E402
from sparsebev_b200 import synthetic as S      # noqa: E402


def main():
    so = os.path.join(ROOT, 'oracle', '_ref', '_msmv_sampling_cuda.so')
    spec = importlib.util.spec_from_file_location('_msmv_sampling_cuda', so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device('cuda:0')
    out = {}
    for name, T in [('r50_704x256', 8), ('r50_704x256', 1)]:
        cfg = S.layer_cfg(name, T, num_layers=1)
        sd = {k: v.to(dev) for k, v in S.make_state_dict(cfg, seed=0).items()}
        feats = [f.to(dev) for f in R.regroup_feats(S.make_feats(name, T, batch=1, seed=100), channel_last=True)]
        metas = S.make_metas(name, T, batch=1)
        td = R.time_diff_from_timestamps([m['img_timestamp'] for m in metas]).to(dev)
        l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas]).astype(np.float32)).to(dev)
        qb = S.init_query_bbox(cfg['num_query'], seed=2)[None].contiguous().to(dev)
        qf = torch.randn(1, cfg['num_query'], 256, generator=torch.Generator().manual_seed(3)).to(dev)
        fwd = ref._ms_deform_attn_cuda_c2345_forward

        def op(mlvl, loc, w):
            return fwd(*mlvl, loc.contiguous(), w.contiguous())

        def step():
            with torch.no_grad():
                return R.decoder_layer(qb, qf, feats, sd, cfg, td, l2i, op=op)
        for _ in range(5):
            step()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        n = 30
        for _ in range(n):
            step()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / n
        out['%s_T%d' % (name, T)] = {'ms_per_layer': ms, 'samples_per_s': 1e3 / ms}
        print(name, T, out['%s_T%d' % (name, T)], flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'gpurun_out', 'ref_layer_gpu.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
