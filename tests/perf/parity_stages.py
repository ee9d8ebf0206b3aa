"""Where does the whole-layer error come from at full size?  Per-stage error of the CUDA layer against the CPU oracle's taps,
each stage fed with the ORACLE's input (so errors do not compound).  Test infrastructure / development tool."""
import copy
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_torch as R                              # noqa: E402
import sparsebev_b200 as sb                                    # noqa: E402
from sparsebev_b200 import ops, synthetic as S                 # noqa: E402


def stats(name, got, want):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    e = (got - want).abs().reshape(-1)
    scale = float(want.abs().max())
    q = torch.quantile(e[torch.randperm(e.numel())[:2000000]], torch.tensor([0.5, 0.99, 0.9999]))
    print('  %-34s max %.3e  p50 %.2e  p99 %.2e  p99.99 %.2e  (of max |ref| = %.3g; ref rms %.3g)' % (
        name, float(e.max()) / scale, float(q[0]) / scale, float(q[1]) / scale, float(q[2]) / scale, scale, float(want.pow(2).mean().sqrt())), flush=True)


def main():
    for name, T in (('r50_704x256', 8), ('vov99_1600x640', 2), ('tiny', 8)):
        cfg = S.layer_cfg(name, T, num_layers=1)
        sd = S.make_state_dict(cfg, seed=1)
        Q = cfg['num_query']
        model = sb.SparseBEVTransformer(256, num_frames=T, num_points=4, num_layers=1, num_levels=cfg['num_levels'], pc_range=cfg['pc_range'])
        model.load_state_dict({'decoder.decoder_layer.' + k: v for k, v in sd.items()})
        model = model.cuda().eval()
        layer = model.decoder.decoder_layer
        feats = S.make_feats(name, T, batch=1, seed=2)
        metas = S.make_metas(name, T, batch=1)
        n = int(np.ceil(np.sqrt(Q))) ** 2
        qb = S.init_query_bbox(n, seed=3)[:Q][None].contiguous()
        qb[..., 8:10] = 0.3 * torch.randn(1, Q, 2, generator=torch.Generator().manual_seed(4))
        qf = torch.randn(1, Q, 256, generator=torch.Generator().manual_seed(5))
        td = R.time_diff_from_timestamps([m['img_timestamp'] for m in metas])
        l2i = torch.from_numpy(np.asarray([m['lidar2img'] for m in metas]).astype(np.float32))
        taps = {}
        with torch.no_grad():
            want = R.decoder_layer(qb, qf, R.regroup_feats(feats, channel_last=True), sd, cfg, td, l2i, op=R.msmv_sampling_kernel_semantics, taps=taps)
            pos = qf + R.position_encoder(qb[..., :3], sd)
        print('== %s T=%d Q=%d' % (name, T, Q), flush=True)
        mg = copy.deepcopy(metas)
        model.decoder.prepare_metas(mg, 1, torch.device('cuda'))
        gf = model.decoder.prepare_feats([f.cuda() for f in feats])
        got = layer(qb.cuda(), qf.cuda(), gf, None, mg)
        for g, w, nm in zip(got, want, ('LAYER query_feat', 'LAYER cls', 'LAYER bbox')):
            stats(nm, g, w)
        # stage by stage, oracle inputs
        M, D = Q, 256
        qbc, x0 = qb.cuda(), qf.cuda().reshape(M, D)
        q1 = torch.empty(M, D, device='cuda')
        ops.dense_chain(qbc.reshape(M, 10), 10, M, [layer._pe0.layer(relu=True), layer._pe1.layer(relu=True, residual=x0, y=q1)])
        stats('pos_enc + query_feat', q1, pos[0])
        sasa = layer.self_attn.forward_fused(qbc, pos.cuda(), None, layer.norm1)
        stats('SASA block (+norm1)', sasa, taps['after_sasa'])
        heads = layer.sampling._heads(taps['after_sasa'].cuda().reshape(M, D))
        G, P, L = 4, 4, cfg['num_levels']
        pts, sw = ops.sample_points(qbc, heads, heads[:, G * P * 3:], cfg['pc_range'], L, num_points_total=G * P, ld_off=heads.shape[1], ld_log=heads.shape[1])
        want_pts = taps['points'][:, :, 0].reshape(1, Q, G * P, 3).clone()          # frame 0: time_diff 0 -> un-warped
        stats('sample points (frame 0)', pts, want_pts)
        stats('scale weights', sw.reshape(1, Q, G, P, L), taps['scale_weights'][:, :, :, 0])
        sampled = layer.sampling(qbc, taps['after_sasa'].cuda(), gf, mg)
        stats('sampled features', sampled, taps['sampled'])
        # sampled features given the ORACLE's points (isolates the gather from the heads Linear)
        sampled2 = ops.sampling4d_fused(gf, want_pts.cuda().contiguous(), qbc, mg[0]['time_diff'], mg[0]['lidar2img'],
                                        taps['scale_weights'][:, :, :, 0].contiguous().cuda(), cfg['image_h'], cfg['image_w'], num_frames=T,
                                        layout=layer.sampling.feat_layout)
        stats('sampled features | oracle points', sampled2, taps['sampled'])
        mixed = layer.mixing.forward_fused(taps['sampled'].cuda(), taps['after_sasa'].cuda(), layer.norm2)
        stats('mixing block (+norm2)', mixed, taps['mixed'])
        with torch.no_grad():
            ffn_want = R._ln(R.ffn(taps['mixed'], sd), sd['norm3.weight'], sd['norm3.bias'])
        q4 = torch.empty(M, D, device='cuda')
        mx = taps['mixed'].cuda().reshape(M, D).contiguous()
        ops.dense_chain(mx, D, M, [layer._ffn0.layer(relu=True), layer._ffn1.layer(residual=mx, res_pre_ln=True, y=q4)])
        stats('FFN (+norm3)', q4, ffn_want[0])
        del model, gf, feats
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
