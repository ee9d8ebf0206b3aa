"""Weights-stationary dense chain (sbev_dense_chain_ws_fwd / _ws_reduce_fwd, csrc/dense_ws.cu) against an fp64 torch
restatement of the reference's Linear / LayerNorm / ReLU / residual chains (/root/reference/models/sparsebev_transformer.py:113-183)
and against the streaming chain kernel, for the five chain shapes of the decoder layer, at the row counts of the 1 / 2 / 4 / 8-GPU
query shards and at ragged ones, with one and two row groups per CTA.  The blob layout itself is checked on CPU (test_ws_blob_layout)."""
import copy

import pytest
import torch

import oracle.ref_torch as R


def dev():
    return torch.device('cuda:0')


def _ops():
    from sparsebev_b200 import ops
    return ops


@pytest.fixture
def option():
    from sparsebev_b200 import _lib
    saved = {}

    def set_(name, value):
        if name not in saved:
            saved[name] = _lib.get_option(name)
        _lib.set_option(name, value)
    yield set_
    for k, v in saved.items():
        _lib.set_option(k, v)


def test_ws_blob_layout():
    """pack_ws_blob: rank c, layer i, [Kpad/64][hi|lo][SW][64 k], row j = feature c*SW + j, 16-byte chunk p of a row holds chunk p ^ (j & 7)."""
    from sparsebev_b200 import ops
    torch.manual_seed(0)
    shapes = [(256, 64), (776, 256), (10, 256), (512, 256), (256, 512)]
    ws = [(torch.randn(n, k).to(torch.bfloat16), torch.randn(n, k).to(torch.bfloat16)) for n, k in shapes]
    blob, stride = ops.pack_ws_blob(ws)
    assert blob.shape == (8, stride) and stride % 128 == 0
    raw = blob.view(torch.bfloat16)                      # [8][stride / 2]
    off = 0
    g = torch.Generator().manual_seed(1)
    for (n, k), (hi, lo) in zip(shapes, ws):
        sw = ops.ws_slice_width(n)
        assert sw % 8 == 0 and 8 * sw >= n
        for _ in range(200):
            c, kc, part, j, p, e = [int(torch.randint(0, m, (1,), generator=g)) for m in (8, k // 64, 2, sw, 8, 8)]
            got = raw[c, off + ((((kc * 2 + part) * sw + j) * 8 + p) * 8) + e]
            f, kk = c * sw + j, kc * 64 + (p ^ (j & 7)) * 8 + e
            want = (hi, lo)[part][f, kk] if f < n else torch.tensor(0., dtype=torch.bfloat16)
            assert got == want, (n, k, c, kc, part, j, p, e)
        off += (k // 64) * 2 * sw * 64
    assert off * 2 <= stride


def test_ws_blob_bytes_match_the_c_side():
    """The host mirror's slice width / blob size arithmetic (ops.ws_slice_width, ops.pack_ws_blob) equals the library's
    (sbev_dense_chain_ws_blob_bytes: host code, callable without a GPU) for the decoder layer's five chains."""
    from sparsebev_b200 import _lib, ops
    lib = _lib.load()
    chains = {'posenc_inproj': [(3, 256), (256, 256), (256, 776)], 'outproj_heads': [(256, 256), (256, 112)], 'ffn': [(256, 512), (512, 256)],
              'cls': [(256, 256), (256, 256), (256, 10)], 'ffn+cls': [(256, 512), (512, 256), (256, 256), (256, 256), (256, 10)]}
    for name, dims in chains.items():
        layers = (_lib.DenseLayer * len(dims))()
        ws = []
        for i, (k, n) in enumerate(dims):
            kpad = (k + 63) // 64 * 64
            layers[i].K, layers[i].N, layers[i].Kpad = k, n, kpad
            ws.append((torch.zeros(n, kpad, dtype=torch.bfloat16), torch.zeros(n, kpad, dtype=torch.bfloat16)))
        want = int(lib.sbev_dense_chain_ws_blob_bytes(len(dims), layers))
        mine = sum((kpad // 64) * 2 * ops.ws_slice_width(n) * 128 for (k, n), (w, _) in zip(dims, ws) for kpad in [w.shape[1]])
        blob, stride = ops.pack_ws_blob(ws)
        assert want == mine and mine <= stride < mine + 128, (name, want, mine, stride)
    assert int(lib.sbev_dense_chain_ws_blob_bytes(0, None)) == -1


def _mk(seed, *dims):
    torch.manual_seed(seed)
    lin = [torch.nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:])]
    return lin


def _ln(n, seed):
    torch.manual_seed(seed)
    ln = torch.nn.LayerNorm(n)
    torch.nn.init.normal_(ln.weight, 1, 0.1)
    torch.nn.init.normal_(ln.bias, 0, 0.1)
    return ln


def _ref(x, specs):
    """fp64 chain: spec = (linear, ln or None, relu, residual or None, res_pre_ln); returns every layer's output."""
    outs = []
    h = x.double()
    for lin, ln, relu, res, pre in specs:
        h = torch.nn.functional.linear(h, lin.weight.double(), lin.bias.double())
        if res is not None and pre:
            h = h + res.double()
        if ln is not None:
            h = torch.nn.functional.layer_norm(h, (h.shape[-1],), ln.weight.double(), ln.bias.double(), 1e-5)
        if relu:
            h = torch.relu(h)
        if res is not None and not pre:
            h = h + res.double()
        outs.append(h)
    return outs


def _entries(ops, specs, ys, extra=None):
    ent, keep = [], []
    for i, (lin, ln, relu, res, pre) in enumerate(specs):
        cache = ops.DenseWeight()
        lind = copy.deepcopy(lin).to(dev())
        lnd = None if ln is None else copy.deepcopy(ln).to(dev())
        wt, ldw, bias = cache.get_with_bias([lind.weight], [lind.bias])
        kw = dict(extra[i]) if extra and extra.get(i) else {}
        ent.append(ops.chain_layer(wt, ldw, lin.in_features, lin.out_features, bias=bias, ln=lnd, relu=relu,
                                   residual=None if res is None else res.to(dev()), res_pre_ln=pre, y=ys[i],
                                   w_hi=cache.w_hi, w_lo=cache.w_lo, kpad=cache.kpad, w_pack=cache.w_pack, **kw))
        keep += [cache, lind, lnd]
    return ent, keep


def _close(got, want, what, rtol=1e-4, atol=1e-4):
    got, want = got.double().cpu(), want.double().cpu()
    err = (got - want).abs()
    ok = err <= atol + rtol * want.abs()
    assert bool(ok.all()), '%s: max abs err %.3e (scale %.3e), %d / %d outside' % (what, float(err.max()), float(want.abs().max()), int((~ok).sum()), ok.numel())


CHAINS = ['posenc_inproj', 'outproj_heads', 'ffn', 'cls', 'reg']


def _build(name, M, seed=0):
    """-> (x, ldx, specs, outputs wanted (indices), refine kwargs or None)"""
    g = torch.Generator().manual_seed(seed + 17)
    rn = lambda *s: torch.randn(*s, generator=g)            # noqa: E731
    if name == 'posenc_inproj':
        l = _mk(seed, 3, 256) + _mk(seed + 1, 256, 256) + _mk(seed + 2, 256, 776)
        x = rn(M, 10)
        return x, 10, [(l[0], _ln(256, 1), True, None, False), (l[1], _ln(256, 2), True, rn(M, 256), False), (l[2], None, False, None, False)], [1, 2], None
    if name == 'outproj_heads':
        l = _mk(seed, 256, 256) + _mk(seed + 1, 256, 112)
        return rn(M, 256), 256, [(l[0], _ln(256, 3), False, rn(M, 256), True), (l[1], None, False, None, False)], [0, 1], None
    if name == 'ffn':
        l = _mk(seed, 256, 512) + _mk(seed + 1, 512, 256)
        x = rn(M, 256)
        return x, 256, [(l[0], None, True, None, False), (l[1], _ln(256, 4), False, x, True)], [1], None
    l = _mk(seed, 256, 256) + _mk(seed + 1, 256, 256) + _mk(seed + 2, 256, 10)
    specs = [(l[0], _ln(256, 5), True, None, False), (l[1], _ln(256, 6), True, None, False), (l[2], None, False, None, False)]
    if name == 'cls':
        return rn(M, 256), 256, specs, [2], None
    qb = R.init_query_bbox(961, seed=2)[:M].contiguous() if M <= 961 else torch.rand(M, 10, generator=g)
    return rn(M, 256), 256, specs, [2], dict(qb=qb, td=torch.tensor([[0.0, 0.5, 1.0]]))


@pytest.mark.gpu
@pytest.mark.parametrize('groups', [0, 1])
@pytest.mark.parametrize('M', [1, 15, 113, 225, 450, 900, 1601])
@pytest.mark.parametrize('name', CHAINS)
def test_ws_chain_vs_fp64(name, M, groups, option):
    if groups != 0 and M not in (450, 900):
        pytest.skip('a forced single row group only where two would be used')
    ops = _ops()
    option('dense_ws', 1)
    option('dense_ws_groups', groups)
    x, ldx, specs, wanted, refine = _build(name, M)
    want = _ref(x[:, :specs[0][0].in_features], specs)
    ys = [torch.full((M, s[0].out_features), float('nan'), device=dev()) if i in wanted else None for i, s in enumerate(specs)]
    extra = {}
    his = {}
    if name in ('posenc_inproj', 'outproj_heads'):          # the layers that also emit the bf16 (hi, lo) operands of the next kernel
        i = 2 if name == 'posenc_inproj' else 0
        his[i] = (torch.zeros(M, specs[i][0].out_features, device=dev(), dtype=torch.bfloat16), torch.zeros(M, specs[i][0].out_features, device=dev(), dtype=torch.bfloat16))
        extra[i] = dict(y_hi=his[i][0], y_lo=his[i][1])
    kw = {}
    if refine is not None:
        extra[2] = dict(refine=True)
        kw = dict(refine_proposal=refine['qb'].to(dev()), refine_time_diff=refine['td'].to(dev()), refine_Q=M, refine_T=3)
        wb = R.refine_bbox(refine['qb'][None], want[2][None].float())
        want[2] = torch.cat([wb[..., :8], wb[..., 8:] / 0.5], -1)[0]
    ent, keep = _entries(ops, specs, ys, extra)
    assert ops.ws_eligible(M, ent), 'the chain should be expressible by the weights-stationary kernel'
    from sparsebev_b200 import _lib
    n0 = _lib.launch_count
    ops.dense_chain(x.to(dev()), ldx, M, ent, **kw)
    torch.cuda.synchronize()
    assert _lib.launch_count == n0 + 1
    for i in wanted:
        _close(ys[i], want[i], '%s M=%d layer %d' % (name, M, i))
    for i, (hi, lo) in his.items():
        _close(hi.float() + lo.float(), want[i], '%s M=%d layer %d (hi + lo)' % (name, M, i), rtol=1e-4, atol=1e-4)
    # and the streaming kernel on the same entries: the two kernels agree to fp32 round-off
    option('dense_ws', 0)
    ys2 = [None if y is None else torch.empty_like(y) for y in ys]
    ent2, keep3 = _entries(ops, specs, ys2, {k: {kk: vv for kk, vv in v.items() if kk == 'refine'} for k, v in extra.items()})
    ops.dense_chain(x.to(dev()), ldx, M, ent2, **kw)
    for i in wanted:
        _close(ys[i], ys2[i], '%s M=%d layer %d vs streaming chain' % (name, M, i), rtol=2e-5, atol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('groups', [0, 1])
@pytest.mark.parametrize('M,nsplit,ln', [(900, 18, True), (113, 128, True), (113, 18, True), (37, 5, False), (1601, 18, True)])
def test_ws_chain_reduce_vs_fp64(M, nsplit, ln, groups, option):
    """sbev_dense_chain_ws_reduce_fwd: input rows = LN(sum_z partial + bias + residual), the FFN behind it."""
    ops = _ops()
    option('dense_ws', 1)
    option('dense_ws_groups', groups)
    torch.manual_seed(5)
    K0 = 256
    part, bias, res = torch.randn(nsplit, M, K0), torch.randn(K0), torch.randn(M, K0)
    norm_in = _ln(K0, 9) if ln else None
    x_in = part.double().sum(0) + bias.double() + res.double()
    if ln:
        x_in = torch.nn.functional.layer_norm(x_in, (K0,), norm_in.weight.double(), norm_in.bias.double(), 1e-5)
    l = _mk(3, 256, 512) + _mk(4, 512, 256)
    x_out = torch.full((M, K0), float('nan'), device=dev())
    specs = [(l[0], None, True, None, False), (l[1], _ln(256, 4), False, x_in.float(), True)]
    want = _ref(x_in, [(l[0], None, True, None, False), (l[1], specs[1][1], False, x_in, True)])
    y = torch.full((M, 256), float('nan'), device=dev())
    ent, keep = _entries(ops, specs, [None, y])
    # the residual of the second layer is the row the prologue itself produces (as in the decoder layer)
    ent[1][0].residual = x_out.data_ptr()
    nd = None if norm_in is None else copy.deepcopy(norm_in).to(dev())
    assert ops.ws_eligible(M, ent, reduce_k0=K0)
    ops.dense_chain_reduce(part.to(dev()), bias.to(dev()), res.to(dev()), None if nd is None else nd.weight, None if nd is None else nd.bias, x_out, ent)
    torch.cuda.synchronize()
    _close(x_out, x_in, 'reduce prologue rows', rtol=1e-5, atol=2e-5 * max(1.0, float(x_in.detach().abs().max())))
    _close(y, want[1], 'ffn behind the reduce prologue')


@pytest.mark.gpu
def test_ws_not_eligible_falls_to_streaming(option):
    """A chain whose weight slices do not fit next to a row tile (FFN + cls: 200 KB per CTA) is declined by ws_eligible and
    runs on the streaming kernel -- same results."""
    ops = _ops()
    option('dense_ws', 1)
    M = 64
    x, ldx, s1, _, _ = _build('ffn', M)
    _, _, s2, _, _ = _build('cls', M, seed=3)
    specs = s1 + s2
    want = _ref(x, specs)
    ys = [None] * 4 + [torch.empty(M, 10, device=dev())]
    ent, keep = _entries(ops, specs, ys)
    assert not ops.ws_eligible(M, ent)
    ops.dense_chain(x.to(dev()), ldx, M, ent)
    _close(ys[4], want[4], 'ffn + cls on the streaming kernel')


@pytest.mark.gpu
@pytest.mark.parametrize('M', [9, 113, 900, 905])
@pytest.mark.parametrize('name', ['cls', 'reg', 'outproj_heads'])
def test_wide_cta_chain_is_bit_identical(name, M, option):
    """SBEV_DENSE_WIDE_CTA (16 rows per CTA of the streaming chain, the form cls || reg use when they run side by side): every row goes
    through the same arithmetic as in the 8-row form -> bit-identical outputs; and both match the fp64 restatement."""
    ops = _ops()
    option('dense_ws', 0)
    x, ldx, specs, wanted, refine = _build(name, M)
    want = _ref(x, specs)
    kw, extra8, extra16 = {}, {}, {0: dict(wide_cta=True)}
    if refine is not None:
        extra8[2] = dict(refine=True)
        extra16[2] = dict(refine=True)
        kw = dict(refine_proposal=refine['qb'].to(dev()), refine_time_diff=refine['td'].to(dev()), refine_Q=M, refine_T=3)
        wb = R.refine_bbox(refine['qb'][None], want[2][None].float())
        want[2] = torch.cat([wb[..., :8], wb[..., 8:] / 0.5], -1)[0]
    outs = []
    for extra in (extra8, extra16):
        ys = [torch.full((M, s[0].out_features), float('nan'), device=dev()) if i in wanted else None for i, s in enumerate(specs)]
        ent, keep = _entries(ops, specs, ys, extra)
        ops.dense_chain(x.to(dev()), ldx, M, ent, **kw)
        torch.cuda.synchronize()
        outs.append(ys)
    for i in wanted:
        _close(outs[1][i], want[i], '%s M=%d layer %d (16-row CTAs)' % (name, M, i))
        assert torch.equal(outs[0][i], outs[1][i]), '%s M=%d layer %d: 16-row CTAs differ from 8-row CTAs' % (name, M, i)
