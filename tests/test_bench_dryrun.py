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
            graph_pool_handle=lambda: 0, is_current_stream_capturing=lambda: False, empty_cache=lambda: None)
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
                                      '--skip-cpu', '--skip-backbone'])
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
    monkeypatch.setattr(ops, 'sampling4d_fused', lambda *a, **k: None)
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
    assert set(d['roofline']) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} and d['roofline']['bound'] == 'hbm'
    assert abs(d['roofline']['frac'] - d['roofline']['achieved'] / d['roofline']['peak']) < 1e-9
    assert 'clocks' in d and 'cpu_baseline' in d
    # e2e = one reference-facing forward (6 decoder layers) per feature upload; the stricter per-layer-upload figure rides along
    e = d['e2e']
    assert d['e2e_decoder_error'] is None, d['e2e_decoder_error']
    assert e['decoder_layer_samples_per_step'] == 6 and e['value'] > 0 and e['unit'] == 'samples/s'
    assert e['h2d_bytes_per_step'] > 0 and e['d2h_bytes_per_step'] == 6 * 2 * 36 * 10 * 4
    assert d['e2e_per_layer_upload']['value'] > 0 and d['e2e_per_layer_upload']['h2d_bytes_per_step'] == e['h2d_bytes_per_step']
    assert d['e2e_resident_features']['value'] > 0 and d['e2e_resident_features']['eager_ms_per_step'] > 0
    assert calls['layer'] > 6 * 10


def test_bench_refuses_to_run_without_cuda(monkeypatch):
    """The real bench (no shims) must fail loudly on a box without CUDA: there is no CPU path for our arm."""
    import pytest
    import bench
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    monkeypatch.setattr(sys, 'argv', ['bench.py', '--config', 'tiny', '--frames', '2', '--steps', '1', '--warmup', '1'])
    with pytest.raises(AssertionError, match='needs a GPU'):
        bench.main()
