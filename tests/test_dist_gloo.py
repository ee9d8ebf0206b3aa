"""N>1 host logic on CPU: world_size-2 gloo process group (rendezvous on 127.0.0.1)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    from sparsebev_b200 import dist as D
    r, w = D.init(backend='gloo')
    assert (r, w) == (rank, world)
    D.barrier()
    mx = D.max_over_ranks(10.0 + rank)
    t0, t1 = D.frame_partition(8, rank, world)
    local = torch.arange(t0, t1, dtype=torch.float32)[:, None].repeat(1, 3)          # [T_local, 3], value = global frame id
    full = D.all_gather_frames(local)
    scenes = D.scene_partition(5, rank, world)
    # frame-sharded exchange: every rank contributes the rows of its frames, the merged tensor is the unsharded one
    B, Q, G, T, P, C = 2, 5, 4, 8, 3, 6
    whole = torch.arange(B * Q * G * T * P * C, dtype=torch.float32).reshape(B, Q, G, T * P, C)
    shard = D.FrameShard(T, exchange='nccl')
    assert shard.window == (t0, t1) and (shard.rank, shard.world) == (rank, world)
    merged = shard.all_gather(whole[:, :, :, t0 * P:t1 * P].contiguous())
    assert torch.equal(merged, whole)
    # partitioning A: feature all-gather, NCHW and channels-last memory, B = 1 (no re-layout) and B = 2
    N, Cc, H, Wd = 6, 4, 3, 5
    for Bf in (1, 2):
        pyramid = torch.arange(Bf * T * N * Cc * H * Wd, dtype=torch.float32).reshape(Bf, T * N, Cc, H, Wd)
        mine = pyramid[:, t0 * N:t1 * N]
        got = D.all_gather_features([mine.contiguous(), mine.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)])
        assert torch.equal(got[0], pyramid) and got[0].is_contiguous()
        assert torch.equal(got[1], pyramid) and got[1].permute(0, 1, 3, 4, 2).is_contiguous()      # stays channels-last
    q.put((rank, mx, full[:, 0].tolist(), scenes))
    torch.distributed.destroy_process_group()


def test_gloo_world2_partitions_and_timing_reduce():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] == 11.0 for r in res)                       # max over ranks
    assert all(r[2] == [0, 1, 2, 3, 4, 5, 6, 7] for r in res)   # frame-major all-gather in rank order
    assert res[0][3] == [0, 1, 2] and res[1][3] == [3, 4]       # scene partition covers everything once


def test_merge_frame_chunks_order():
    from sparsebev_b200 import dist as D
    W, B, Q, G, Tl, P, C = 4, 1, 3, 4, 2, 4, 2
    whole = torch.randn(B, Q, G, W * Tl * P, C)
    chunks = torch.stack([whole[:, :, :, r * Tl * P:(r + 1) * Tl * P] for r in range(W)])
    assert torch.equal(D.merge_frame_chunks(chunks), whole)
    one = D.FrameShard(8, rank=0, world=1)
    assert one.window == (0, 8) and one.all_gather(whole) is whole


def test_partitions_single_process():
    from sparsebev_b200 import dist as D
    assert D.frame_partition(8, 3, 4) == (6, 8)
    assert sum((D.scene_partition(7, r, 3) for r in range(3)), []) == list(range(7))
    assert D.max_over_ranks(3.5) == 3.5
    x = torch.ones(2, 4)
    assert D.all_gather_frames(x) is x
    f = [torch.ones(1, 6, 4, 2, 2)]
    assert D.all_gather_features(f)[0] is f[0]                    # no process group: identity


def test_query_shard_partition_layout_and_segments(monkeypatch):
    """Host logic of the query- and frame-sharded decoder (no GPU, no process group: explicit rank / world): the query
    partition covers every query once (ragged tail included), arena fields start on 256-byte boundaries behind the flag
    words, and an exchange hands sbev_peer_exchange the own rows at the SAME offset of every rank's buffer."""
    from sparsebev_b200 import dist as D, ops
    for Q, world in ((900, 8), (900, 4), (900, 2), (1600, 8), (36, 8), (5, 8), (40, 3)):
        seen = []
        for r in range(world):
            qpr, q0, q1 = D.query_partition(Q, r, world)
            assert qpr * world >= Q and 0 <= q0 <= q1 <= Q and q1 - q0 <= qpr
            seen += list(range(q0, q1))
            if q1 > q0:
                assert q0 // qpr == r and (q1 - 1) // qpr == r          # the gather's owner rule: q // q_per_rank
        assert seen == list(range(Q))
    fields = [('points', (900, 16, 3)), ('scale_w', (900, 16, 4)), ('sampled', (113, 4, 32, 64)), ('box0', (900, 10))]
    table, total = D.QueryShard.layout(fields)
    prev_end = D.QueryShard.FLAG_WORDS
    for name, shape in fields:
        off, shp = table[name]
        assert shp == shape and off % 64 == 0 and off >= prev_end
        prev_end = off + D._numel(shape)
    assert total >= prev_end
    sh = D.QueryShard(8, rank=3, world=8)
    assert sh.window == (3, 4) and sh.partition(900) == (113, 339, 452)
    calls = []
    monkeypatch.setattr(ops, 'peer_exchange', lambda segs, n, r, flags, ctl, dev: calls.append((segs, n, r, flags)))
    bases = [1 << 40 | (w << 32) for w in range(8)]
    ar = dict(ptrs=bases, table=table, ctl=torch.zeros(4, dtype=torch.int32), device='cpu')
    sh.exchange(ar, [('points', 339, 452), ('box0', 339, 452)])
    sh.exchange(ar, [])
    (segs, n, r, flags), (segs2, _, _, _) = calls
    assert (n, r, flags) == (8, 3, bases) and segs2 == [] and sh.exchanges == 2
    for (src, dsts, nbytes), (name, per) in zip(segs, (('points', 48), ('box0', 10))):
        start = 4 * (table[name][0] + 339 * per)
        assert src == bases[3] + start and dsts == [b + start for b in bases] and nbytes == 4 * 113 * per
    assert sh.bytes_sent == 7 * 4 * 113 * (48 + 10)
