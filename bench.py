#!/usr/bin/env python
"""bench.py -- decoder-layer samples/s of the SparseBEV hot path on B200 (BASELINE.json metric).

One STEP = one scene (B=1: 900 queries x 6 cameras x T=8 frames, r50 704x256 FPN, 4 levels) pushed through ONE
decoder layer (position encoding -> scale-adaptive self-attention -> adaptive spatio-temporal sampling ->
adaptive mixing -> FFN -> cls/reg heads -> box refinement), i.e. SparseBEVTransformerDecoderLayer.forward.

  python bench.py --gpus N --steps K --warmup W           our arm  (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                    reference arm: the reference's native-PyTorch CPU path
                                                          (oracle port; /root/reference does not exist on the GPU box)
Prints ONE JSON line (rank 0).  Keys: see the task contract; extra keys are documented in DESIGN.md.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

METRIC = 'decoder-layer samples/sec (900q x 6cam x 8f)'
GATHER_DRAM_TRAFFIC = 131.6e6      # bytes per launch at r50-T8, measured with ncu --set full (profiles/r01_kernels_ncu.md)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='r50_704x256')
    ap.add_argument('--frames', type=int, default=8)
    ap.add_argument('--precision', default='bf16x3', choices=['bf16x3', 'bf16'])
    ap.add_argument('--layout', default='grouped', choices=['grouped', 'nhwc'])
    ap.add_argument('--no-graph', action='store_true')
    ap.add_argument('--no-overlap', action='store_true', help='single stream: no concurrent gather || param-GEMM, cls || reg')
    ap.add_argument('--no-tma-params', action='store_true', help='mixing: fp32 parameter tensor + converting mix kernel instead of bf16 (hi,lo) + TMA')
    ap.add_argument('--opt', action='append', default=[], metavar='NAME=VALUE', help='kernel-variant option passed to sbev_set_option (experiments)')
    ap.add_argument('--split-k', type=int, default=None, help='split-K slices of the mixing output projection')
    ap.add_argument('--shard', default='scenes', choices=['scenes', 'frames'],
                    help='N>1: scenes = one scene per GPU (weak scaling, no collective; default); frames = ONE scene, every GPU holds and samples '
                         'T/N frames and the sampled rows are exchanged once per layer (strong scaling)')
    ap.add_argument('--exchange', default='p2p', choices=['p2p', 'nccl'], help='--shard frames: peer stores from the gather kernel, or NCCL all-gather')
    ap.add_argument('--breakdown', action='store_true', help='also write per-stage timings to gpurun_out/breakdown.json')
    ap.add_argument('--cpu-steps', type=int, default=3, help='bounded CPU sample: decoder-layer passes of the oracle')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--skip-backbone', action='store_true', help='do not time the ResNet-50 + FPN image branch (SURVEY 8 a17) next to the headline metric')
    return ap.parse_args()


# --------------------------------------------------------------------------------------------- helpers
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[4:8]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def gather_algorithmic_bytes(cfg, B=1, G=4, C=64):
    """per sampled point read L*4 corners*C*4 B of features + (3+L)*4 B of coords/weights, write C*4 B."""
    L, T, P, Q = cfg['num_levels'], cfg['num_frames'], cfg['num_points'], cfg['num_query']
    points = B * T * G * Q * P
    return points * (L * 4 * C * 4 + (3 + L) * 4 + C * 4), points


# ------------------------------------------------------------------------------------------ CPU oracle arm
def cpu_layer_timer(cfg, steps, threads=None):
    """The reference's native-PyTorch decoder-layer path restated in oracle/ref_torch.py (F.grid_sample sampling,
    eager mixing / attention), fp32, on the host CPU.  The thread count is auto-tuned (one probe step each for
    8/16/32/64/all cores; eager PyTorch on small tensors gets SLOWER when oversubscribed) so the baseline gets its
    best shot.  Returns (best seconds, mean seconds, threads used)."""
    from oracle import ref_torch as R
    from sparsebev_b200 import synthetic as S
    ncpu = os.cpu_count() or 1
    T = cfg['num_frames']
    sd = S.make_state_dict(cfg, seed=0)
    feats = R.regroup_feats(S.make_feats(cfg['name'], T, batch=1, seed=1), channel_last=False)
    metas = S.make_metas(cfg['name'], T, batch=1)
    td = R.time_diff_from_timestamps([m['img_timestamp'] for m in metas])
    l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas]).astype(np.float32))
    qb = S.init_query_bbox(cfg['num_query'], seed=2)[None].contiguous()
    qf = torch.randn(1, cfg['num_query'], 256, generator=torch.Generator().manual_seed(3))
    def one():
        t0 = time.perf_counter()
        with torch.no_grad():
            R.decoder_layer(qb, qf, feats, sd, cfg, td, l2i)
        return time.perf_counter() - t0

    if threads is None:
        best_t, best = None, None
        for t in sorted(set(min(c, ncpu) for c in (8, 16, 32, 64, ncpu))):
            torch.set_num_threads(t)
            one()
            dt = one()
            if best is None or dt < best:
                best_t, best = t, dt
        threads = best_t
    torch.set_num_threads(threads)
    one()
    times = [one() for _ in range(steps)]
    return min(times), float(np.mean(times)), threads


def run_reference_arm(args, cfg):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    best, mean, threads = cpu_layer_timer(cfg, steps)
    val = 1.0 / mean
    print(json.dumps({
        'metric': METRIC, 'value': val, 'unit': 'samples/s', 'n_gpus': args.gpus, 'steps': steps, 'warmup': 1,
        'ms_per_step': mean * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'impl': 'reference',
        'config': {'workload': '%s T=%d Q=%d L=%d, one decoder layer, B=1' % (args.config, cfg['num_frames'], cfg['num_query'], cfg['num_levels'])},
        'cpu_baseline': {'value': val, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                         'sample': '%d decoder-layer passes (mean; best %.1f ms) of the reference native-PyTorch path '
                                   'restated in oracle/ref_torch.py; %d of %d host threads (auto-tuned)' % (steps, best * 1e3, threads, os.cpu_count() or 1)},
        'e2e': {'value': val, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def _shutdown(graph=None):
    """Tear the process group down; a watchdog ends the process if NCCL teardown stalls (e.g. a captured graph still
    holding communicator work) so a finished benchmark can never sit on the GPU box until the caller's timeout."""
    import torch.distributed as dist
    sys.stdout.flush()
    threading.Timer(30.0, lambda: os._exit(0)).start()
    del graph
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    os._exit(0)


# ------------------------------------------------------------------------------------------------- ours
def main():
    args = parse()
    from sparsebev_b200 import synthetic as S
    cfg = S.layer_cfg(args.config, args.frames, num_layers=1)
    cfg['name'] = args.config
    if args.impl == 'reference':
        return run_reference_arm(args, cfg)

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py (our arm) needs a GPU; there is no CPU fallback'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    import sparsebev_b200 as sb
    from sparsebev_b200 import _lib, ops

    T, Q = cfg['num_frames'], cfg['num_query']
    model = sb.SparseBEVTransformer(256, num_frames=T, num_points=cfg['num_points'], num_layers=1,
                                    num_levels=cfg['num_levels'], pc_range=cfg['pc_range'])
    model.load_state_dict({'decoder.decoder_layer.' + k: v for k, v in S.make_state_dict(cfg, seed=0).items()})
    model = model.to(dev).eval()
    layer = model.decoder.decoder_layer
    layer.mixing.precision = args.precision
    layer.overlap = not args.no_overlap
    layer.mixing.tma_params = not args.no_tma_params
    if args.split_k:
        layer.mixing.split_k = args.split_k
    for kv in args.opt:
        name, value = kv.split('=')
        _lib.set_option(name, int(value))

    frames_mode = args.shard == 'frames' and world > 1
    if frames_mode:
        # strong scaling: ONE scene; rank r holds the feature maps of frames [r*T/N, (r+1)*T/N) only
        from sparsebev_b200 import dist as D
        shard = D.FrameShard(T, exchange=args.exchange)
        model.shard_frames(shard)
        t0, t1 = shard.window
        feats_host = [f[:, t0 * 6:t1 * 6].contiguous() for f in
                      S.make_feats(args.config, T, batch=1, seed=100, memory_format='nhwc' if args.layout == 'nhwc' else 'nchw')]
    else:
        # weak scaling: every rank owns its own scene (different seed) -- the reference's only strategy is DP
        feats_host = S.make_feats(args.config, T, batch=1, seed=100 + rank, memory_format='nhwc' if args.layout == 'nhwc' else 'nchw')
    metas = S.make_metas(args.config, T, batch=1)
    model.decoder.prepare_metas(metas, 1, dev)
    feats = model.decoder.prepare_feats([f.to(dev) for f in feats_host])
    qb_host = S.init_query_bbox(Q, seed=2)[None].contiguous().pin_memory()
    qf_host = torch.randn(1, Q, 256, generator=torch.Generator().manual_seed(3)).pin_memory()
    qb, qf = qb_host.to(dev), qf_host.to(dev)
    feat_bytes = sum(f.numel() * 4 for f in feats)

    def step():
        return layer(qb, qf, feats, None, metas)

    # ---- warm-up (also builds weight caches / sets function attributes), count launches of one step
    torch.cuda.synchronize()
    for _ in range(max(args.warmup - 1, 2)):
        step()                                   # first call also builds the per-weight device caches (one-time launches)
    n0 = _lib.launch_count
    step()
    launches_per_step = _lib.launch_count - n0     # steady state: kernels of OURS per decoder-layer pass
    torch.cuda.synchronize()

    graph = None
    if not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step()
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(graph):
            outs = step()
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()

    def run_step():
        if graph is not None:
            graph.replay()
        else:
            step()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    # ---- timed region: exactly K steps, device-timed, max over ranks
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        run_step()
    ev1.record()
    torch.cuda.synchronize(); barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([elapsed_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = world * 1.0 / (ms_per_step * 1e-3)
    if frames_mode:
        clk = clocks.stop() if rank == 0 else None
        if rank == 0:
            print(json.dumps({
                'metric': METRIC, 'value': 1.0 / (ms_per_step * 1e-3), 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
                'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': '%s T=%d Q=%d L=%d, one decoder layer, ONE scene (B=1) across %d GPUs' % (args.config, T, Q, cfg['num_levels'], world),
                           'tensor_core_precision': '%s on tcgen05 / mma.sync, fp32 accumulate' % args.precision,
                           'l2': 'inputs larger than L2 (feature pyramid %.0f MB per GPU per step)' % (feat_bytes / 1e6),
                           'parallelism': 'frame-sharded: %d frames per GPU, sampled rows exchanged once per layer (%s)' % (T // world, args.exchange),
                           'cuda_graph': graph is not None, 'two_stream_overlap': layer.overlap},
                'clocks': clk, 'gpu_launches': launches_per_step * args.steps, 'launches_per_step': launches_per_step}))
        _shutdown(graph)
        return

    # ---- e2e: public API call with HOST buffers; every step copies that step's inputs (query tensors, camera
    # metadata AND the feature pyramid) from pinned host memory and reads the results back.  Feature upload of step
    # i+1 is double-buffered on a copy stream so it overlaps the compute of step i.
    pinned_feats = [f.contiguous().cpu().pin_memory() for f in feats]
    dbuf = [[torch.empty_like(f) for f in feats] for _ in range(2)]
    l2i_host = metas[0]['lidar2img'].cpu().pin_memory()
    td_host = metas[0]['time_diff'].cpu().pin_memory()
    out_host = [torch.empty(1, Q, 256).pin_memory(), torch.empty(1, Q, 10).pin_memory(), torch.empty(1, Q, 10).pin_memory()]
    copy_stream = torch.cuda.Stream()
    h2d = feat_bytes + qb_host.numel() * 4 + qf_host.numel() * 4 + l2i_host.numel() * 4 + td_host.numel() * 4
    d2h = sum(o.numel() * 4 for o in out_host)
    e2e_steps = max(3, min(args.steps, 10))

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            for d, s in zip(dbuf[slot], pinned_feats):
                d.copy_(s, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return ev

    def e2e_loop(n):
        ready = upload(0)
        for i in range(n):
            slot = i & 1
            torch.cuda.current_stream().wait_event(ready)
            if i + 1 < n:
                ready = upload(slot ^ 1)
            m = [dict(metas[0])]
            m[0]['lidar2img'] = l2i_host.to(dev, non_blocking=True)
            m[0]['time_diff'] = td_host.to(dev, non_blocking=True)
            o = layer(qb_host.to(dev, non_blocking=True), qf_host.to(dev, non_blocking=True), dbuf[slot], None, m)
            for h, d in zip(out_host, o):
                h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()

    e2e_loop(2)
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    # ---- e2e through the reference-facing call: ONE SparseBEVTransformer-style forward per scene = upload the scene's
    # inputs once, run the 6 shared-weight decoder layers on them (reference: num_layers=6, configs/r50_nuimg_704x256.py:70-79;
    # layer i+1 consumes layer i's boxes and features, sparsebev_transformer.py:86-99), read the stacked predictions back.
    # That is 6 decoder-layer samples per call, so the rate in layer-samples/s is 6 x calls/s.  Guarded: if anything in this
    # newer loop fails, the per-layer-upload figure above is reported as `e2e` instead and the error is recorded.
    NUM_DEC_LAYERS = 6
    e2e_dec_ms, e2e_dec_err, d2h_dec = None, None, 0
    try:
        dec_host = [[torch.empty(1, Q, 10).pin_memory(), torch.empty(1, Q, 10).pin_memory()] for _ in range(NUM_DEC_LAYERS)]
        d2h_dec = sum(t.numel() * 4 for pair in dec_host for t in pair)

        def upload_ordered(slot):
            copy_stream.wait_stream(torch.cuda.current_stream())      # the previous user of this slot has been enqueued: order after it
            return upload(slot)

        def e2e_decoder_loop(n):
            ready = upload_ordered(0)
            for i in range(n):
                slot = i & 1
                torch.cuda.current_stream().wait_event(ready)
                if i + 1 < n:
                    ready = upload_ordered(slot ^ 1)
                m = [dict(metas[0])]
                m[0]['lidar2img'] = l2i_host.to(dev, non_blocking=True)
                m[0]['time_diff'] = td_host.to(dev, non_blocking=True)
                qb_d, qf_d = qb_host.to(dev, non_blocking=True), qf_host.to(dev, non_blocking=True)
                for li in range(NUM_DEC_LAYERS):
                    qf_d, cls_d, box_d = layer(qb_d, qf_d, dbuf[slot], None, m)
                    qb_d = box_d
                    dec_host[li][0].copy_(cls_d, non_blocking=True)
                    dec_host[li][1].copy_(box_d, non_blocking=True)
            torch.cuda.synchronize()

        e2e_decoder_loop(2)
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_decoder_loop(e2e_steps)
        e2e_dec_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([e2e_dec_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_dec_ms = float(t.item())
    except Exception as exc:                      # pragma: no cover
        e2e_dec_ms, e2e_dec_err = None, repr(exc)[:300]
    del pinned_feats, dbuf

    # second e2e figure: the reference's real flow -- the pyramid is produced on-device by the backbone, only the query
    # tensors + camera metadata come from the host each step, results are read back
    def e2e_resident(n):
        for _ in range(n):
            m = [dict(metas[0])]
            if layer.use_cuda_graph:       # the graphed layer copies host (pinned) sources straight into its static inputs
                m[0]['lidar2img'], m[0]['time_diff'] = l2i_host, td_host
                o = layer(qb_host, qf_host, feats, None, m)
            else:
                m[0]['lidar2img'] = l2i_host.to(dev, non_blocking=True)
                m[0]['time_diff'] = td_host.to(dev, non_blocking=True)
                o = layer(qb_host.to(dev, non_blocking=True), qf_host.to(dev, non_blocking=True), feats, None, m)
            for hbuf, d in zip(out_host, o):
                hbuf.copy_(d, non_blocking=True)
        torch.cuda.synchronize()

    def time_resident():
        e2e_resident(3)
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_resident(args.steps)
        return (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_res_eager_ms = time_resident()
    e2e_res_ms = e2e_res_eager_ms
    if not args.no_graph:                  # public-API graph mode (layer.use_cuda_graph): one graph launch + the small copies per call
        layer.use_cuda_graph = True
        e2e_res_ms = time_resident()
        layer.use_cuda_graph = False
        layer.reset_graphs()
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([e2e_res_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_res_ms = float(t.item())

    # ---- roofline of the dominant kernel (fused gather), timed alone with CUDA events on the launch stream
    x = qf.reshape(Q, 256)
    heads = layer.sampling._heads(x)
    GP = 4 * cfg['num_points']
    pts, sw = ops.sample_points(qb, heads, heads[:, GP * 3:], cfg['pc_range'], cfg['num_levels'], num_points_total=GP,
                                ld_off=heads.shape[1], ld_log=heads.shape[1])
    vel = qb[..., 8:10].contiguous()
    sw5 = sw.reshape(1, Q, 4, cfg['num_points'], cfg['num_levels'])
    out_buf = torch.empty(1, Q, 4, T * cfg['num_points'], 64, device=dev)

    def gather():
        ops.sampling4d_fused(feats, pts, vel, metas[0]['time_diff'], metas[0]['lidar2img'], sw5, cfg['image_h'], cfg['image_w'],
                             num_frames=T, layout=layer.sampling.feat_layout, out=out_buf)
    for _ in range(3):
        gather()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_g = 20
    torch.cuda.synchronize()
    g0.record()
    for _ in range(n_g):
        gather()
    g1.record()
    torch.cuda.synchronize()
    gather_ms = g0.elapsed_time(g1) / n_g
    algo_bytes, n_points = gather_algorithmic_bytes(cfg)
    peak, peak_src = measured_peaks()
    achieved = algo_bytes / (gather_ms * 1e-3) / 1e9

    # ---- second roofline: the tensor-core GEMM of the mixing stage (dynamic-parameter generation), timed alone the same way
    mixing = layer.mixing
    pbuf = mixing.alloc_params(Q, dev)
    mixing.generate_params(x, pbuf, presplit=False)          # fills the bf16 (hi, lo) query operand once
    for _ in range(3):
        mixing.generate_params(x, pbuf, presplit=True)
    torch.cuda.synchronize()
    g0.record()
    for _ in range(n_g):
        mixing.generate_params(x, pbuf, presplit=True)
    g1.record()
    torch.cuda.synchronize()
    gemm_ms = g0.elapsed_time(g1) / n_g
    n_par = mixing.n_groups * mixing.total_parameters
    products = 3 if args.precision == 'bf16x3' else 1
    gemm_flops = 2.0 * Q * n_par * 256 * products            # tensor-core flops issued (bf16x3 = three bf16 products per fp32-grade one)
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            tensor_peak, tensor_src = float(json.load(f)['bf16_tflops']), 'measured (MEASURED_PEAKS.json bf16_tflops, burst)'
    except Exception:
        tensor_peak, tensor_src = 2250.0, 'fallback (nominal dense bf16)'
    gemm_tflops = gemm_flops / (gemm_ms * 1e-3) / 1e12

    clk = clocks.stop() if rank == 0 else None

    breakdown = None
    if args.breakdown and rank == 0:
        breakdown = stage_breakdown(layer, qb, qf, feats, metas, cfg)
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', 'breakdown.json'), 'w') as f:
            json.dump(breakdown, f, indent=1)

    backbone = None
    if rank == 0 and world == 1 and not args.skip_backbone:
        backbone = backbone_record(dev, cfg)

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        best, mean, threads = cpu_layer_timer(cfg, args.cpu_steps)
        cpu = {'value': 1.0 / mean, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
               'sample': '%d decoder-layer passes (mean %.0f ms, best %.0f ms) of the reference native-PyTorch path '
                         '(oracle/ref_torch.py), same workload; %d of %d host threads (auto-tuned)' % (args.cpu_steps, mean * 1e3, best * 1e3, threads, os.cpu_count() or 1)}

    per_layer_e2e = {'value': world * 1e3 / e2e_ms, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                     'ms_per_step': e2e_ms, 'steps': e2e_steps,
                     'note': 'ONE decoder layer per step with all its inputs incl. the %.0f MB feature pyramid uploaded from pinned host memory '
                             'every step (double-buffered on a copy stream), results read back: PCIe-bound by construction' % (feat_bytes / 1e6)}
    if rank == 0:
        print(json.dumps({
            'metric': METRIC, 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': '%s T=%d Q=%d L=%d, one decoder layer, B=1 per GPU' % (args.config, T, Q, cfg['num_levels']),
                       'l2': 'inputs larger than L2 (feature pyramid %.0f MB per step)' % (feat_bytes / 1e6),
                       'tensor_core_precision': '%s on tcgen05 / mma.sync, fp32 accumulate' % args.precision,
                       'feat_layout': layer.sampling.feat_layout, 'cuda_graph': graph is not None, 'two_stream_overlap': layer.overlap, 'parallelism': 'dp%d (one scene per GPU)' % world},
            'clocks': clk,
            'e2e': per_layer_e2e if e2e_dec_ms is None else
                   {'value': world * NUM_DEC_LAYERS * 1e3 / e2e_dec_ms, 'unit': 'samples/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h_dec,
                    'ms_per_step': e2e_dec_ms, 'steps': e2e_steps, 'decoder_layer_samples_per_step': NUM_DEC_LAYERS,
                    'note': 'one reference-facing forward per step: ALL its inputs (query tensors, camera metadata, the %.0f MB feature pyramid) '
                            'uploaded from pinned host memory (double-buffered on a copy stream), the %d shared-weight decoder layers run on them '
                            '(eager launches), every layer\'s cls / bbox predictions read back; value = %d decoder-layer samples per forward / time'
                            % (feat_bytes / 1e6, NUM_DEC_LAYERS, NUM_DEC_LAYERS)},
            'e2e_per_layer_upload': per_layer_e2e,
            'e2e_decoder_error': e2e_dec_err,
            'e2e_resident_features': {'value': world * 1e3 / e2e_res_ms, 'unit': 'samples/s', 'ms_per_step': e2e_res_ms,
                                      'h2d_bytes_per_step': h2d - feat_bytes, 'd2h_bytes_per_step': d2h,
                                      'eager_ms_per_step': e2e_res_eager_ms,
                                      'note': 'public API call (layer.use_cuda_graph: the layer replays its own captured graph; eager_ms_per_step = plain launches); '
                                              'feature pyramid already on the device as in the reference pipeline'},
            'gpu_launches': launches_per_step * args.steps,
            'launches_per_step': launches_per_step,
            'roofline': {'bound': 'hbm', 'kernel': 'sampling4d_c64_kernel (fused projection + multi-scale gather)',
                         'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': GATHER_DRAM_TRAFFIC,
                         'traffic_note': 'dram__bytes_read.sum + dram__bytes_write.sum of this kernel from ncu --set full (profiles/); far below '
                                         'the algorithmic bytes because neighbouring queries re-sample the same pixels out of L2',
                         'algorithmic_bytes': algo_bytes, 'points': n_points, 'kernel_ms': gather_ms, 'peak_source': peak_src,
                         'dram_achieved': GATHER_DRAM_TRAFFIC / (gather_ms * 1e-3) / 1e9, 'dram_frac': GATHER_DRAM_TRAFFIC / (gather_ms * 1e-3) / 1e9 / peak},
            'roofline_tensor': {'bound': 'tensor', 'kernel': 'gemm_bf16_tn_persistent_kernel (mixing parameter generation, [%d x 256] x [256 x %d], %s)' % (Q, n_par, args.precision),
                                'achieved': gemm_tflops, 'peak': tensor_peak, 'unit': 'TFLOP/s', 'frac': gemm_tflops / tensor_peak,
                                'kernel_ms': gemm_ms, 'flops': gemm_flops, 'fp32_grade_tflops': gemm_tflops / products, 'peak_source': tensor_src,
                                'note': 'bf16x3 issues three bf16 tensor-core products per fp32-grade product; the kernel is bound by L2->SM operand '
                                        'traffic (~390 MB per launch), not by the tensor pipe'},
            'cpu_baseline': cpu,
            'backbone': backbone,
            'breakdown_ms': breakdown}))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def backbone_record(dev, cfg, images=6):
    """Secondary record (BASELINE config 4's image branch): ResNet-50 + FPN over the 6 camera images of one new
    frame (what the reference's online inference runs per sample, models/sparsebev.py:255-321), our tcgen05 implicit-GEMM
    convolutions (NHWC bf16, fp32 FPN levels in the gather's layout) vs the same modules through stock PyTorch / cuDNN
    (bf16 autocast, channels_last).  Device-timed with CUDA events after warm-up; random-init weights."""
    import torch.nn.functional as F
    from sparsebev_b200 import backbone as BB
    try:
        torch.manual_seed(0)
        net = BB.ResNet(depth=50).to(dev).eval()
        neck = BB.FPN([256, 512, 1024, 2048], 256, cfg['num_levels']).to(dev).eval()
        img = torch.randn(1, images, 3, cfg['image_h'], cfg['image_w'], device=dev)

        def timeit(fn, iters=10):
            for _ in range(3):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / iters

        def stock(x):
            with torch.autocast('cuda', dtype=torch.bfloat16):
                x = net.maxpool(net.relu(net.bn1(net.conv1(x))))
                outs = []
                for name in net.res_layers:
                    for blk in getattr(net, name):
                        idt = x if blk.downsample is None else blk.downsample(x)
                        o = blk.relu(blk.bn2(blk.conv2(blk.relu(blk.bn1(blk.conv1(x))))))
                        x = blk.relu(blk.bn3(blk.conv3(o)) + idt)
                    outs.append(x)
                lats = [l.conv(f) for l, f in zip(neck.lateral_convs, outs)]
                for i in range(len(lats) - 1, 0, -1):
                    lats[i - 1] = lats[i - 1] + F.interpolate(lats[i], size=lats[i - 1].shape[-2:], mode='nearest')
                return [c.conv(l).float() for c, l in zip(neck.fpn_convs, lats)]
        def graphed(fn):
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            return timeit(g.replay)
        with torch.no_grad():
            ours = timeit(lambda: BB.extract_img_feat(net, neck, img))
            ours_graph = graphed(lambda: BB.extract_img_feat(net, neck, img))
            flat = img[0].contiguous(memory_format=torch.channels_last)
            net.to(memory_format=torch.channels_last); neck.to(memory_format=torch.channels_last)
            ref = timeit(lambda: stock(flat))
            try:
                ref_graph = graphed(lambda: stock(flat))
            except Exception:
                ref_graph = None
        return {'workload': 'ResNet-50 + FPN (%d levels), %d images %dx%d' % (cfg['num_levels'], images, cfg['image_w'], cfg['image_h']),
                'ms': ours_graph, 'images_per_s': images * 1e3 / ours_graph, 'eager_ms': ours,
                'stock_pytorch_cudnn_bf16_ms': ref_graph, 'stock_pytorch_cudnn_bf16_eager_ms': ref, 'dtype': 'bf16 operands, fp32 accumulate',
                'note': 'ms = CUDA-graph replay (63 launches of ours: sbev_stem_conv_fwd, sbev_maxpool3x3s2_nhwc_fwd, sbev_conv2d_nhwc_fwd x 61); '
                        'the stock PyTorch arm runs the same modules through cuDNN under bf16 autocast, channels_last'}
    except Exception as e:                                    # secondary record: never take the headline line down with it
        return {'error': '%s: %s' % (type(e).__name__, e)}


def stage_breakdown(layer, qb, qf, feats, metas, cfg, iters=20):
    """Per-stage device time (CUDA events, eager launches) -- explains ms_per_step; not part of the metric."""
    from sparsebev_b200 import ops
    B, Q, D = qf.shape
    M = B * Q
    res = {}

    def timeit(name, fn):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            r = fn()
        b.record()
        torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / iters
        return r

    x0 = qf.reshape(M, D).contiguous()
    def pos_enc():
        q1 = torch.empty(M, D, device=qf.device)
        ops.dense_chain(qb.reshape(M, -1), qb.shape[-1], M, [layer._pe0.layer(relu=True), layer._pe1.layer(relu=True, residual=x0, y=q1)])
        return q1
    h = timeit('pos_enc_chain', pos_enc)
    q1 = timeit('sasa_block', lambda: layer.self_attn.forward_fused(qb, h.reshape(B, Q, D), None, layer.norm1))
    sampled = timeit('sampling_block', lambda: layer.sampling(qb, q1, feats, metas))
    q2 = timeit('mixing_block', lambda: layer.mixing.forward_fused(sampled, q1, layer.norm2))
    x2 = q2.reshape(M, D)

    def ffn_cls():
        q4, cls = torch.empty(M, D, device=qf.device), torch.empty(M, layer.num_classes, device=qf.device)
        ch = [layer._ffn0.layer(relu=True), layer._ffn1.layer(residual=x2, res_pre_ln=True, y=q4)]
        ch += [l.layer(relu=True) for l in layer._cls[:-1]] + [layer._cls[-1].layer(y=cls)]
        ops.dense_chain(x2, D, M, ch)
        return q4
    q3 = timeit('ffn_cls_chain', ffn_cls)

    def reg():
        box = torch.empty(M, 10, device=qf.device)
        td = metas[0]['time_diff']
        ops.dense_chain(q3, D, M, [l.layer(relu=True) for l in layer._reg[:-1]] + [layer._reg[-1].layer(refine=True, y=box)],
                        refine_proposal=qb, refine_time_diff=td, refine_Q=Q, refine_T=td.shape[1])
    timeit('reg_refine_chain', reg)
    # finer split of the mixing block
    mix = layer.mixing
    qh, ql = ops.split_bf16(q1.reshape(M, D))
    wh, wl = mix._pg.get(mix.parameter_generator.weight)
    npar = mix.n_groups * mix.total_parameters
    seg_a, seg_b = ([qh, qh, ql], [wh, wl, wh]) if mix.precision == 'bf16x3' else ([qh], [wh])
    params = timeit('mix.param_gemm', lambda: ops.gemm_bf16_tn(seg_a, seg_b, M, npar, D, bias=mix.parameter_generator.bias))
    yh, yl, _ = timeit('mix.mix_kernel', lambda: ops.mix(params, sampled.reshape(M, 4, -1, 64)))
    oh, ol = mix._op.get(mix.out_proj.weight)
    sa, sb_ = ([yh, yh, yl], [oh, ol, oh]) if mix.precision == 'bf16x3' else ([yh], [oh])
    part = timeit('mix.out_gemm', lambda: ops.gemm_bf16_tn(sa, sb_, M, D, mix.out_proj.in_features, split_k=mix.split_k))
    timeit('mix.reduce_ln', lambda: ops.reduce_ln(part, mix.out_proj.bias, q1.reshape(M, D), layer.norm2.weight, layer.norm2.bias))
    # SASA split
    qkvt = torch.empty(M, 3 * D + 8, device=qf.device)
    timeit('sasa.in_proj_tau', lambda: ops.dense_chain(h, D, M, [layer.self_attn.in_layer(qkvt)]))
    timeit('sasa.core(split+v3)', lambda: ops.sasa_split(qkvt, qb, cfg['pc_range'], 8, D))
    timeit('sasa.core(v2)', lambda: ops.sasa(qkvt, qb, qkvt[:, 3 * D:], cfg['pc_range'], 8, ld_qkv=3 * D + 8, ld_tau=3 * D + 8, embed_dims=D))
    heads = layer.sampling._heads(q1.reshape(M, D))
    timeit('sampling.gather', lambda: layer.sampling.sample(qb, heads, feats, metas))
    return res


if __name__ == '__main__':
    main()
