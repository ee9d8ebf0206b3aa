"""2+-rank GPU check of the frame-sharded decoder (launched by tests/test_gpu_multi.py under torchrun).

Every rank builds the SAME synthetic scene, runs the unsharded decoder on all frames, then the frame-sharded decoder on
its own window of the feature maps with both exchanges (peer stores over NVLink, NCCL all-gather).  The gather computes
every sample independently of the window, and everything else is replicated, so the outputs must be bit-identical.
Prints one line per rank: `FRAME_SHARD_OK rank=<r> exchange=<e> ...`; any mismatch raises.
"""
import copy
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import sparsebev_b200 as sb                    # noqa: E402
from sparsebev_b200 import dist as D           # noqa: E402
from sparsebev_b200 import synthetic as S      # noqa: E402


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else 'tiny'
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    exchanges = sys.argv[3].split(',') if len(sys.argv) > 3 else ['p2p', 'nccl']
    rank, world, local = D.env_rank_world()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    D.init('nccl', dev)
    cfg = S.layer_cfg(config, T, num_layers=2)
    sd = S.make_state_dict(cfg, seed=0)
    model = sb.SparseBEVTransformer(embed_dims=256, num_frames=T, num_points=4, num_layers=2, num_levels=cfg['num_levels'],
                                    pc_range=cfg['pc_range']).to(dev).eval()
    model.load_state_dict({'decoder.decoder_layer.' + k: v for k, v in sd.items()})
    B, Q = 1, cfg['num_query']
    feats = S.make_feats(config, T, batch=B, seed=1)
    metas = S.make_metas(config, T, batch=B)
    qb = S.init_query_bbox(Q, seed=2)[None].repeat(B, 1, 1).to(dev)
    qf = torch.randn(B, Q, 256, generator=torch.Generator().manual_seed(3)).to(dev)
    with torch.no_grad():
        want = model(qb, qf, [f.to(dev) for f in feats], None, copy.deepcopy(metas))
        t0, t1 = D.frame_partition(T, rank, world)
        for ex in exchanges:
            shard = D.FrameShard(T, exchange=ex)
            assert shard.window == (t0, t1)
            model.shard_frames(shard)
            local_feats = [f[:, t0 * 6:t1 * 6].contiguous().to(dev) for f in feats]
            for rep in range(3):          # repeated forwards also exercise the alternating peer buffers
                got = model(qb, qf, [f.clone() for f in local_feats], None, copy.deepcopy(metas))
                torch.cuda.synchronize()
                for g, w in zip(got, want):
                    if not torch.equal(g, w):
                        raise AssertionError('rank %d exchange %s rep %d: max |diff| %.3e' % (rank, ex, rep, float((g - w).abs().max())))
            print('FRAME_SHARD_OK rank=%d world=%d exchange=%s window=[%d,%d) config=%s' % (rank, world, ex, t0, t1, config), flush=True)
        model.shard_frames(None)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    try:
        main()
    except BaseException:
        import traceback
        print('FRAME_SHARD_FAILED rank=%s\n%s' % (os.environ.get('RANK'), traceback.format_exc()), flush=True)
        os._exit(1)                    # do not hang the peers in a collective teardown
