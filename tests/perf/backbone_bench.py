"""Backbone conv path timing on one B200: our tcgen05 implicit-GEMM ResNet-50 + FPN (NHWC bf16) against the same modules run
by stock PyTorch (cuDNN, bf16 autocast, channels_last) -- the path the reference's conv wrappers take (F.conv2d).
Prints one JSON line; per-conv-shape table with --layers.  usage: python tests/perf/backbone_bench.py [--images 6] [--layers]"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from sparsebev_b200 import backbone as BB, ops  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def torch_forward(net, neck, img):
    with torch.autocast('cuda', dtype=torch.bfloat16):
        x = net.maxpool(net.relu(net.bn1(net.conv1(img))))
        outs = []
        for name in net.res_layers:
            for blk in getattr(net, name):
                idt = x if blk.downsample is None else blk.downsample(x)
                o = blk.relu(blk.bn1(blk.conv1(x)))
                o = blk.relu(blk.bn2(blk.conv2(o)))
                x = blk.relu(blk.bn3(blk.conv3(o)) + idt)
            outs.append(x)
        lats = [l.conv(f) for l, f in zip(neck.lateral_convs, outs)]
        for i in range(len(lats) - 1, 0, -1):
            lats[i - 1] = lats[i - 1] + F.interpolate(lats[i], size=lats[i - 1].shape[-2:], mode='nearest')
        res = [c.conv(l) for c, l in zip(neck.fpn_convs, lats)]
    return [r.float() for r in res]


def conv_flops(net, neck, H, W):
    """2 * MACs of every conv for one image."""
    total = 0
    shapes = []
    h, w = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    total += 2 * 147 * 64 * h * w
    h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    for name in net.res_layers:
        for blk in getattr(net, name):
            s = blk.conv2.stride[0]
            for conv, (hh, ww, st) in ((blk.conv1, (h, w, 1)), (blk.conv2, (h, w, s)), (blk.conv3, (h // s, w // s, 1))):
                ho, wo = hh // st, ww // st
                total += 2 * conv.in_channels * conv.out_channels * conv.kernel_size[0] ** 2 * ho * wo
                shapes.append((conv.in_channels, conv.out_channels, conv.kernel_size[0], st, hh, ww))
            if blk.downsample is not None:
                d = blk.downsample[0]
                total += 2 * d.in_channels * d.out_channels * (h // s) * (w // s)
                shapes.append((d.in_channels, d.out_channels, 1, s, h, w))
            h, w = h // s, w // s
    hh, ww = (H // 4, W // 4)
    for i, (l, c) in enumerate(zip(neck.lateral_convs, neck.fpn_convs)):
        total += 2 * l.conv.in_channels * 256 * hh * ww + 2 * 256 * 256 * 9 * hh * ww
        shapes.append((l.conv.in_channels, 256, 1, 1, hh, ww)); shapes.append((256, 256, 3, 1, hh, ww))
        hh, ww = hh // 2, ww // 2
    return total, shapes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--images', type=int, default=6)
    ap.add_argument('--height', type=int, default=256)
    ap.add_argument('--width', type=int, default=704)
    ap.add_argument('--layers', action='store_true')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    net = BB.ResNet(depth=50).to(dev).eval()
    neck = BB.FPN([256, 512, 1024, 2048], 256, 4).to(dev).eval()
    img = torch.randn(1, args.images, 3, args.height, args.width, device=dev)
    flat = img.reshape(args.images, 3, args.height, args.width)
    with torch.no_grad():
        ours = timeit(lambda: BB.extract_img_feat(net, neck, img))
        net_cl, neck_cl = net.to(memory_format=torch.channels_last), neck.to(memory_format=torch.channels_last)
        flat_cl = flat.contiguous(memory_format=torch.channels_last)
        torch.backends.cudnn.benchmark = True
        ref = timeit(lambda: torch_forward(net_cl, neck_cl, flat_cl))
        a = BB.extract_img_feat(net, neck, img)
        b = torch_forward(net_cl, neck_cl, flat_cl)
        err = max(float((x.reshape(y.shape) - y).abs().max() / y.abs().max()) for x, y in zip(a, b))
    flops, shapes = conv_flops(net, neck, args.height, args.width)
    out = {'workload': 'ResNet-50 + FPN, %d images %dx%d' % (args.images, args.width, args.height), 'ours_ms': ours, 'torch_cudnn_bf16_ms': ref,
           'speedup': ref / ours, 'gflop': flops * args.images / 1e9, 'ours_tflops': flops * args.images / ours / 1e9,
           'torch_tflops': flops * args.images / ref / 1e9, 'rel_diff_vs_torch_bf16': err}
    if args.layers:
        rows = []
        for (cin, cout, k, s, h, w) in sorted(set(shapes)):
            x = torch.randn(args.images, h, w, cin, device=dev).bfloat16()
            wt = torch.randn(cout, k, k, cin, device=dev).bfloat16()
            sh = torch.zeros(cout, device=dev)
            t = timeit(lambda: ops.conv2d_nhwc(x, wt, sh, None, stride=s, pad=k // 2, relu=True), iters=20)
            xc = x.permute(0, 3, 1, 2)
            wc = wt.permute(0, 3, 1, 2)
            tr = timeit(lambda: F.relu(F.conv2d(xc, wc, None, s, k // 2)), iters=20)
            ho, wo = (h + 2 * (k // 2) - k) // s + 1, (w + 2 * (k // 2) - k) // s + 1
            fl = 2.0 * cin * cout * k * k * ho * wo * args.images
            rows.append({'cin': cin, 'cout': cout, 'k': k, 's': s, 'h': h, 'w': w, 'ours_us': t * 1e3, 'cudnn_us': tr * 1e3,
                         'ours_tflops': fl / t / 1e9, 'cudnn_tflops': fl / tr / 1e9, 'count': shapes.count((cin, cout, k, s, h, w))})
        out['layers'] = rows
    print(json.dumps(out))


if __name__ == '__main__':
    main()
