"""Evaluation step starting from the classification head's NCHW logits, three ways, on one B200:
  (a) reference ops for the head tail (torch sigmoid + permute) then criterion(...) + decoder(...)
  (b) b200det.heads.sigmoid_channels_last then b200det.fused.EvalStep
  (c) b200det.fused.LogitsEvalStep (no probability tensor)
RetinaNet-R50 800x800 head shapes, CUDA events, inputs far larger than L2.

    python tools/bench_logits.py [--batch 256] [--dtype f32|f16|bf16]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from b200det import heads, synth, losses, decode, fused, _lib  # noqa: E402


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
    ap.add_argument('--dtype', default='f32')
    ap.add_argument('--iters', type=int, default=20)
    args = ap.parse_args()
    dtype = {'f32': torch.float32, 'f16': torch.float16, 'bf16': torch.bfloat16}[args.dtype]
    B, C, A = args.batch, 80, 9
    gen = torch.Generator(device='cuda').manual_seed(0)
    cls, reg = [], []
    for p in synth.pyramid_sizes(800):
        cls.append((torch.randn((B, A * C, p, p), generator=gen, device='cuda') - 4.595).to(dtype))
        reg.append(torch.randn((B, p, p, A, 4), generator=gen, device='cuda') * 0.2)
    ann = synth.make_annotations(B, 100, 800, C, seed=1).cuda()
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    dec = decode.RetinaDecoder(**synth.RETINA_KW)
    estep = fused.EvalStep(crit, dec)
    lstep = fused.LogitsEvalStep(crit, dec)

    def way_a():
        with torch.no_grad():
            probs = [torch.sigmoid(x.float()).permute(0, 2, 3, 1).contiguous()
                     .view(B, x.shape[2], x.shape[3], A, C) for x in cls]
            return crit([probs, reg], ann), dec([probs, reg])

    def way_b():
        with torch.no_grad():
            probs = [heads.sigmoid_channels_last(x, C) for x in cls]
            return estep([probs, reg], ann)

    def way_c():
        with torch.no_grad():
            return lstep([cls, reg], ann)

    ta, tb = timed(way_a, max(args.iters // 2, 3)), timed(way_b, args.iters)
    tc = timed(way_c, args.iters)
    _lib.profile_start()
    for _ in range(5):
        way_c()
    torch.cuda.synchronize()
    prof = _lib.profile_stop()
    n = sum(x.numel() for x in cls)
    peak = 6549.1
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs']
    except Exception:
        pass
    sweep_ms = prof['logits_sweep'][1]
    print(json.dumps({
        'workload': f'eval step from NCHW logits, batch {B}, RetinaNet-R50 800^2, {args.dtype}',
        'ms': {'torch_tail+criterion+decoder': round(ta, 3), 'fused_tail+EvalStep': round(tb, 3),
               'LogitsEvalStep': round(tc, 3)},
        'images_per_s': {'torch_tail+criterion+decoder': round(B / ta * 1e3),
                         'fused_tail+EvalStep': round(B / tb * 1e3), 'LogitsEvalStep': round(B / tc * 1e3)},
        'logits_sweep_ms': round(sweep_ms, 4),
        'logits_sweep_GBps': round(n * cls[0].element_size() / sweep_ms / 1e6, 1),
        'logits_sweep_frac_of_hbm_peak': round(n * cls[0].element_size() / sweep_ms / 1e6 / peak, 3),
        'kernels_ms': {k: round(v[1], 4) for k, v in prof.items()},
    }))


if __name__ == '__main__':
    main()
