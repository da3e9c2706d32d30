"""Head tail (b200det.heads) vs the reference's torch ops on one B200: RetinaNet-R50 800x800 cls
head outputs (720 channels x 5 levels), CUDA events, inputs far larger than L2.

    python tools/bench_heads.py [--batch 256] [--dtype f32|f16|bf16]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from b200det import heads, synth, _lib  # noqa: E402


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--channels', type=int, default=720)
    ap.add_argument('--size', type=int, default=800)
    ap.add_argument('--dtype', default='f32')
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    dtype = {'f32': torch.float32, 'f16': torch.float16, 'bf16': torch.bfloat16}[args.dtype]
    sizes = synth.pyramid_sizes(args.size)
    gen = torch.Generator(device='cuda').manual_seed(0)
    xs = [(torch.randn((args.batch, args.channels, p, p), generator=gen, device='cuda') - 4.6).to(dtype)
          for p in sizes]
    n = sum(x.numel() for x in xs)
    in_b = xs[0].element_size()

    def ref_fwd():
        return [torch.sigmoid(x.float()).permute(0, 2, 3, 1).contiguous() for x in xs]

    def our_fwd():
        return [heads.sigmoid_channels_last(x) for x in xs]

    t_ref = timed(ref_fwd, args.iters)
    t_our = timed(our_fwd, args.iters)
    ys = our_fwd()
    gs = [torch.randn_like(y) for y in ys]

    def ref_bwd():
        out = []
        for x, y, g in zip(xs, ys, gs):
            gy = g.permute(0, 3, 1, 2)                       # permute backward (view)
            out.append((gy * (1 - y.permute(0, 3, 1, 2)) * y.permute(0, 3, 1, 2)).to(dtype).contiguous())
        return out

    def our_bwd():
        out = []
        for x, y, g in zip(xs, ys, gs):
            b, h, w, c = y.shape
            gi = torch.empty((b, c, h, w), dtype=dtype, device='cuda')
            _lib.check(_lib.load().b200det_head_sigmoid_permute_backward(
                g.data_ptr(), y.data_ptr(), b, c, h * w, gi.data_ptr(), heads._DTYPES[dtype],
                heads._stream()), 'bwd')
            out.append(gi)
        return out

    t_rb = timed(ref_bwd, max(args.iters // 2, 2))
    t_ob = timed(our_bwd, args.iters)
    peak = 6549.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    fwd_bytes = n * (in_b + 4)
    bwd_bytes = n * (8 + in_b)
    print(json.dumps({
        'workload': f'cls head tail, batch {args.batch}, {args.channels} ch, {args.size}^2, {args.dtype}',
        'elements': n,
        'fwd_ms': {'torch_ops': round(t_ref, 3), 'b200det': round(t_our, 3)},
        'fwd_GBps': round(fwd_bytes / t_our / 1e6, 1), 'fwd_frac_of_hbm_peak': round(fwd_bytes / t_our / 1e6 / peak, 3),
        'bwd_ms': {'torch_ops': round(t_rb, 3), 'b200det': round(t_ob, 3)},
        'bwd_GBps': round(bwd_bytes / t_ob / 1e6, 1), 'bwd_frac_of_hbm_peak': round(bwd_bytes / t_ob / 1e6 / peak, 3),
    }))


if __name__ == '__main__':
    main()
