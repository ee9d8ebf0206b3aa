"""Experiment (DESIGN 6.1 / VERDICT r1 #4): does running parameter GEMM -> mix -> out-projection GEMM over ROW CHUNKS keep the
dynamic parameters and the mixed tile L2-resident (126 MB L2) instead of round-tripping 470 MB through HBM?
Same kernels, same results; only the launch granularity changes.  CUDA-graph replays, device-timed.  Writes
gpurun_out/mix_chunk_bench.json.  Test infrastructure."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sparsebev_b200 import ops  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    M, D, G, Pin, C = 900, 256, 4, 32, 64
    NP, K2 = G * (C * C + 128 * Pin), G * 128 * C
    g = torch.Generator().manual_seed(0)
    q = torch.randn(M, D, generator=g).to(dev)
    W = (torch.randn(NP, D, generator=g) * 0.02).to(dev)
    b = (torch.randn(NP, generator=g) * 0.05).to(dev)
    Wo = (torch.randn(D, K2, generator=g) * 0.01).to(dev)
    x = torch.randn(M, G, Pin, C, generator=g).to(dev)
    qh, ql = ops.split_bf16(q)
    wh, wl = ops.split_bf16(W)
    oh, ol = ops.split_bf16(Wo)
    res = {}
    ref_out = None
    for chunk, two_stream in ((900, False), (450, False), (256, False), (128, False), (256, True), (128, True)):
        nbuf = 2 if two_stream else 1
        ph = [torch.empty(chunk, NP, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        pl = [torch.empty_like(ph[0]) for _ in range(nbuf)]
        yh = [torch.empty(chunk, K2, device=dev, dtype=torch.bfloat16) for _ in range(nbuf)]
        yl = [torch.empty_like(yh[0]) for _ in range(nbuf)]
        out = torch.empty(M, D, device=dev)
        rows = [(r0, min(M, r0 + chunk)) for r0 in range(0, M, chunk)]
        units = max(1, (min(chunk, M) + 255) // 256)
        split_k = max(18, min(128, (72 + units - 1) // units))
        parts = [torch.empty(split_k, r1 - r0, D, device=dev) for r0, r1 in rows]
        side = torch.cuda.Stream()

        def run():
            main_s = torch.cuda.current_stream()
            evs = []
            for i, (r0, r1) in enumerate(rows):
                n, k = r1 - r0, i % nbuf
                if two_stream:           # parameter GEMM of chunk i on the side stream, overlapping mix / out-projection of chunk i-1
                    side.wait_stream(main_s) if i < nbuf else side.wait_event(evs[i - nbuf])
                    with torch.cuda.stream(side):
                        ops.gemm_bf16_tn_split(qh[r0:r1], ql[r0:r1], wh, wl, n, NP, D, bias=b, out=(ph[k][:n], pl[k][:n]))
                    main_s.wait_stream(side)
                else:
                    ops.gemm_bf16_tn_split(qh[r0:r1], ql[r0:r1], wh, wl, n, NP, D, bias=b, out=(ph[k][:n], pl[k][:n]))
                ops.mix_presplit(ph[k][:n], pl[k][:n], x[r0:r1], out=(yh[k][:n], yl[k][:n]))
                ev = torch.cuda.Event(); ev.record(main_s); evs.append(ev)
                ops.gemm_bf16_tn([yh[k][:n], yh[k][:n], yl[k][:n]], [oh, ol, oh], n, D, K2, split_k=split_k, out=parts[i])
                out[r0:r1] = ops.reduce_ln(parts[i], residual=q[r0:r1])
        run()
        torch.cuda.synchronize()
        if ref_out is None:
            ref_out = out.clone()
        err = float((out - ref_out).abs().max() / ref_out.abs().max())
        gr = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            run()
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(gr):
            run()
        for _ in range(3):
            gr.replay()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(20):
            gr.replay()
        e.record()
        torch.cuda.synchronize()
        key = 'chunk%d%s' % (chunk, '_2stream' if two_stream else '')
        res[key] = {'ms': a.elapsed_time(e) / 20, 'launches': 4 * len(rows), 'split_k': split_k, 'rel_diff_vs_unchunked': err}
        print(key, res[key], flush=True)
        del gr
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'mix_chunk_bench.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
