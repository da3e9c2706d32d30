#!/usr/bin/env python
"""Times every BASELINE.json config (not only the headline one) on one GPU and prints one JSON
object: per config the ms and images/s of loss forward, loss forward+backward and decode+NMS, with
the fraction of the measured HBM peak implied by SURVEY.md section 8(d)'s algorithmic bytes.

    python tools/bench_configs.py [--reps 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from b200det import synth, losses, decode  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        return 6650.0


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(name, kind, size, C, B, G, reps, box='GIoU', sigma=1.0):
    dev = torch.device('cuda')
    sizes = synth.pyramid_sizes(size)
    if kind == 'retina':
        preds = synth.make_retina_preds(B, size, C, seed=1, sigma=sigma, device=dev)
        crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type=box)
        dec = decode.RetinaDecoder(**synth.RETINA_KW)
        N, k = sum(p * p * 9 for p in sizes), 0
    else:
        preds = synth.make_fcos_preds(B, size, C, seed=1, sigma=sigma, device=dev)
        crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
        dec = decode.FCOSDecoder(strides=synth.STRIDES)
        N, k = sum(p * p for p in sizes), 1
    ann = synth.make_annotations(B, G, size, C, seed=2).to(dev)
    loss_b = 4 * N * C + 16 * N + 4 * N * k + 20 * G
    bwd_b = 4 * N * C + 16 * N + 4 * N * k
    dec_b = 4 * N * C + 16 * N + 4 * N * k + 2400

    def fwd():
        with torch.no_grad():
            return crit(preds, ann)

    req = [[t.detach().requires_grad_(True) for t in grp] for grp in preds]

    def fwd_bwd():
        for grp in req:
            for t in grp:
                t.grad = None
        d = crit(req, ann)
        sum(d.values()).backward()

    def dec_fn():
        return dec(preds)

    pk = peak()
    out = {'config': name, 'batch': B, 'rows_per_image': N, 'classes': C}
    for key, fn, nbytes in (('loss_fwd', fwd, loss_b), ('loss_fwd_bwd', fwd_bwd, loss_b + bwd_b),
                            ('decode_nms', dec_fn, dec_b)):
        ms = timed(fn, reps)
        out[key] = {'ms': round(ms, 4), 'images_per_s': round(B / ms * 1e3, 1),
                    'frac_of_hbm_peak': round(B * nbytes / (ms * 1e-3) / 1e9 / pk, 4)}
    del preds, req
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=20)
    args = ap.parse_args()
    res = [
        run('cfg1 Retina 800 C80 B1 (decode named)', 'retina', 800, 80, 1, 100, args.reps),
        run('cfg2 Retina 800 C80 B16 G100', 'retina', 800, 80, 16, 100, args.reps),
        run('cfg2 Retina 800 C80 B16 SmoothL1', 'retina', 800, 80, 16, 100, args.reps, box='SmoothL1'),
        run('cfg3 FCOS 800 C80 B16', 'fcos', 800, 80, 16, 100, args.reps),
        run('cfg4 FCOS 1024 C365 B32 G200', 'fcos', 1024, 365, 32, 200, args.reps),
        run('cfg5 Retina 800 C80 B64 (training-size shard)', 'retina', 800, 80, 64, 100, args.reps),
        # the "sparse" score distribution (sigma 0.5: ~4 % of the rows above 0.05
        # instead of ~98 %) -- fewer candidates for the select kernel, same sweeps
        run('cfg5 Retina 800 C80 B64, sparse scores (sigma 0.5)', 'retina', 800, 80, 64, 100,
            args.reps, sigma=0.5),
        run('cfg3 FCOS 800 C80 B16, sparse scores (sigma 0.5)', 'fcos', 800, 80, 16, 100,
            args.reps, sigma=0.5),
    ]
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
