"""bench.py control flow + JSON contract, dry-run on the CPU.

The GPU is replaced by shims IN THIS TEST ONLY (bench.py itself refuses to run without CUDA and has no CPU path): a fake
`torch.cuda` namespace (streams / events / graphs that do nothing), `torch.device('cuda', i)` mapped to the CPU, pinned-memory
and kernel front-ends (`layer.forward`, `_Dense.__call__`, `ops.sample_points`, `ops.sampling4d_fused`, `generate_params`)
replaced by shape-correct stand-ins.  What is checked is the host logic a typo would break on the GPU box where nobody is
watching: every loop of main() runs, exactly one JSON line comes out, and it carries the keys / types the driver reads.
"""
import contextlib
import io
import json
import sys
import types

import torch


class _Stream:
    cuda_stream = 0

    def wait_stream(self, other): pass
    def wait_event(self, ev): pass
    def synchronize(self): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False


class _Event:
    def __init__(self, enable_timing=False): pass
    def record(self, stream=None): pass
    def elapsed_time(self, other): return 1.0
    def synchronize(self): pass


class _Graph:
    def replay(self): pass


class _TorchProxy(types.ModuleType):
    """`torch` as bench.py sees it: real torch, except device('cuda', i) -> cpu and a do-nothing torch.cuda."""

    def __init__(self):
        super().__init__('torch')
        cuda = types.SimpleNamespace(
            is_available=lambda: True, set_device=lambda i: None, synchronize=lambda *a: None,
            Stream=lambda *a, **k: _Stream(), Event=_Event, current_stream=lambda *a: _Stream(),
            stream=lambda s: contextlib.nullcontext(), CUDAGraph=_Graph, graph=lambda g, **k: contextlib.nullcontext(),
            graph_pool_handle=lambda: 0, is_current_stream_capturing=lambda: False, empty_cache=lambda: None, _sleep=lambda n: None)
        self.__dict__['cuda'] = cuda

    def device(self, *a, **k):
        return torch.device('cpu')

    def __getattr__(self, name):
        return getattr(torch, name)


def test_bench_main_dry_run_emits_one_contract_line(monkeypatch):
    import bench
    from sparsebev_b200 import _lib, ops, transformer as TR
    monkeypatch.setattr(bench, 'torch', _TorchProxy())
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self: self, raising=False)
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--config', 'tiny', '--frames', '2', '--steps', '4', '--warmup', '3',
                                      '--skip-cpu', '--skip-backbone', '--skip-gpu-baseline'])
    for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'):
        monkeypatch.delenv(k, raising=False)
    calls = {'layer': 0}

    def fake_layer(self, query_bbox, query_feat, mlvl_feats, attn_mask, img_metas):
        calls['layer'] += 1
        _lib.launch_count += 11
        assert query_bbox.shape[-1] == 10 and query_feat.shape[-1] == 256 and 'time_diff' in img_metas[0]
        B, Q = query_bbox.shape[:2]
        return torch.zeros(B, Q, 256), torch.zeros(B, Q, 10), torch.full((B, Q, 10), 0.5)
    monkeypatch.setattr(TR.SparseBEVTransformerDecoderLayer, 'forward', fake_layer)
    monkeypatch.setattr(TR._Dense, '__call__', lambda self, x, **k: torch.zeros(x.shape[0], self.out_features))
    monkeypatch.setattr(ops, 'sample_points', lambda qb, off, log, pc, L, **k: (torch.zeros(1), torch.zeros(qb.shape[1] * 4 * 4 * L)))
    def fake_gather(feats, pts, vel, td, l2i, sw, h, w, num_frames, return_loc=False, **k):
        loc = torch.rand(num_frames * 4, 36, 4, 3, generator=torch.Generator().manual_seed(0))
        return (None, loc) if return_loc else None

    def fake_indices(level_hw, loc, num_views):
        Bp, Q, P, _ = loc.shape
        L = len(level_hw)
        view = (loc[..., 2] * (num_views - 1)).round().to(torch.int32)
        y0 = torch.stack([(loc[..., 1] * (h - 1)).floor() for h, w in level_hw], -1).to(torch.int32)
        x0 = torch.stack([(loc[..., 0] * (w - 1)).floor() for h, w in level_hw], -1).to(torch.int32)
        return view, y0, x0, torch.ones(Bp, Q, P, L, dtype=torch.int32)
    monkeypatch.setattr(ops, 'sampling4d_fused', fake_gather)
    monkeypatch.setattr(ops, 'msmv_indices', fake_indices)
    monkeypatch.setattr(ops, 'msmv_forward', lambda feats, loc, w: None)
    monkeypatch.setattr(TR.AdaptiveMixing, 'generate_params', lambda self, q2, buf, presplit=False: None)
    monkeypatch.setattr(_lib, 'set_option', lambda *a: None)

    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        bench.main()
    lines = [l for l in out.getvalue().splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['metric'].startswith('decoder-layer samples/sec') and d['unit'] == 'samples/s' and d['higher_is_better'] is True
    assert d['n_gpus'] == 1 and d['steps'] == 4 and d['warmup'] == 3 and d['scaling'] == 'weak' and d['vs_baseline'] is None
    assert d['dtype'] == 'f32' and d['data'] == 'synthetic' and 'workload' in d['config'] and 'l2' in d['config']
    assert d['value'] > 0 and d['ms_per_step'] > 0 and d['gpu_launches'] == 11 * 4 and d['launches_per_step'] == 11
    assert set(d['roofline']) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic', 'bytes', 'live_tap_fraction'} and d['roofline']['bound'] == 'hbm'
    assert abs(d['roofline']['frac'] - d['roofline']['achieved'] / d['roofline']['peak']) < 1e-9
    assert d['roofline']['traffic'] is None                       # no committed ncu capture of the `tiny` workload: never a made-up constant
    assert 0 < d['roofline']['bytes'] <= d['roofline']['algorithmic_bytes_upper_bound']
    assert d['roofline_uniform']['bound'] == 'hbm' and d['roofline_tensor']['bound'] == 'tensor'
    assert 'clocks' in d and 'cpu_baseline' in d and d['gpu_baseline'] is None
    # e2e = one reference-facing forward (6 decoder layers) per feature upload, host metas converted inside the timed region
    e = d['e2e']
    assert 'error' not in e, e
    assert e['decoder_layer_samples_per_step'] == 6 and e['value'] > 0 and e['unit'] == 'samples/s'
    assert e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] == 6 * 2 * 36 * 10 * 4
    assert d['e2e_resident_features']['value'] > 0 and d['e2e_resident_features']['eager_ms_per_step'] > 0
    assert calls['layer'] > 6 * 5
    # both arms describe the workload with the same config object
    import argparse
    cfg = bench.load_synthetic().layer_cfg('tiny', 2, num_layers=1)
    assert d['config'] == bench.make_config(argparse.Namespace(config='tiny', gpus=1, shard='queries'), cfg)


def test_reference_arm_line_and_hygiene(monkeypatch):
    """`bench.py --impl reference`: honours --steps / --warmup, emits the same `config` object as our arm, and never maps
    the product library into its process (it loads synthetic.py stand-alone, not through the package __init__)."""
    import subprocess
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json, io, contextlib; sys.argv=['bench.py','--impl','reference','--config','tiny','--frames','2','--steps','2','--warmup','1'];"
            "import bench; out=io.StringIO();\n"
            "with contextlib.redirect_stdout(out): bench.main()\n"
            "d=json.loads(out.getvalue().strip().splitlines()[-1]);"
            "maps=open('/proc/self/maps').read();"
            "print(json.dumps({'line': d, 'product_so_mapped': 'libsparsebev_b200' in maps, 'pkg_imported': 'sparsebev_b200' in sys.modules}))")
    r = subprocess.run([sys.executable, '-c', code], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    d = res['line']
    assert res['product_so_mapped'] is False and res['pkg_imported'] is False
    assert d['impl'] == 'reference' and d['steps'] == 2 and d['warmup'] == 1 and d['gpu_launches'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['value'] == d['value'] and d['e2e']['value'] == d['value']
    import argparse
    import bench
    cfg = bench.load_synthetic().layer_cfg('tiny', 2, num_layers=1)
    assert d['config'] == bench.make_config(argparse.Namespace(config='tiny', gpus=1, shard='queries'), cfg)


def test_bench_refuses_to_run_without_cuda(monkeypatch):
    """The real bench (no shims) must fail loudly on a box without CUDA: there is no CPU path for our arm."""
    import pytest
    import bench
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--config', 'tiny', '--frames', '2', '--steps', '1', '--warmup', '1'])
    with pytest.raises(AssertionError, match='needs a GPU'):
        bench.main()


def test_layer_graph_cache_policy_dry_run(monkeypatch):
    """Host policy of the layer's CUDA-graph mode without a GPU (capture / replay are shims; the numerical equality of graph
    and eager launches is a GPU test): one capture per input signature, replays reuse it, a moved feature map or a bumped
    parameter version captures again, the cache keeps MAX_GRAPHS entries, outputs are copies of the static outputs."""
    import sparsebev_b200 as sb
    from sparsebev_b200 import synthetic as S, transformer as TR
    captures = {'n': 0, 'replays': 0}

    class G(_Graph):
        def replay(self):
            captures['replays'] += 1
    fake_cuda = types.SimpleNamespace(Stream=lambda *a, **k: _Stream(), current_stream=lambda *a: _Stream(), stream=lambda s: contextlib.nullcontext(),
                                      CUDAGraph=G, graph=lambda g, **k: contextlib.nullcontext(), graph_pool_handle=lambda: 0,
                                      is_current_stream_capturing=lambda: False)
    monkeypatch.setattr(TR.torch, 'cuda', fake_cuda)
    cfg = S.layer_cfg('tiny', 2, num_layers=1)
    model = sb.SparseBEVTransformer(256, num_frames=2, num_points=4, num_layers=1, num_levels=2, pc_range=cfg['pc_range']).eval()
    layer = model.decoder.decoder_layer

    def impl(qb, qf, feats, mask, metas):
        captures['n'] += 1
        return qf * 2, qb[..., :10] + 1, qb + metas[0]['time_diff'].sum()
    monkeypatch.setattr(layer, '_forward_impl', impl)
    layer.use_cuda_graph = True
    feats = [torch.zeros(1, 12, 256, 8, 22), torch.zeros(1, 12, 256, 4, 11)]
    metas = [dict(img_shape=[(64, 176, 3)], time_diff=torch.ones(1, 2), lidar2img=torch.zeros(1, 12, 4, 4))]
    qb, qf = torch.rand(1, 36, 10), torch.rand(1, 36, 256)
    a = layer(qb, qf, feats, None, metas)
    assert captures['n'] == 2 and captures['replays'] == 1 and len(layer._graphs) == 1          # warm-up pass + capture pass
    static = next(iter(layer._graphs.values()))[1]
    assert torch.equal(static['qb'], qb) and torch.equal(static['td'], metas[0]['time_diff'])
    assert a[0].data_ptr() != next(iter(layer._graphs.values()))[2][0].data_ptr()                 # outputs are copies
    layer(qb * 0.5, qf, feats, None, metas)
    assert captures['n'] == 2 and captures['replays'] == 2 and torch.equal(static['qb'], qb * 0.5)
    keep = [[f.clone() for f in feats]]                                                          # (kept alive: distinct addresses)
    layer(qb, qf, keep[-1], None, metas)                                                         # moved features
    assert len(layer._graphs) == 2
    with torch.no_grad():
        layer.norm1.weight.add_(1.0)                                                             # parameter version bump
    layer(qb, qf, feats, None, metas)
    assert len(layer._graphs) == 3
    for _ in range(4):
        keep.append([f.clone() for f in feats])
        layer(qb, qf, keep[-1], None, metas)
    assert len(layer._graphs) == layer.MAX_GRAPHS
    layer.reset_graphs()
    assert len(layer._graphs) == 0
