"""Per-kernel device time vs wall time of one decoder / criterion call at small batches
(BASELINE configs 0-3 are latency-bound).  python tools/prof_small.py"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from b200det import synth, losses, decode, _lib  # noqa: E402

out = {}
for name, B, fcos in (('retina_b1', 1, False), ('retina_b16', 16, False), ('fcos_b16', 16, True)):
    if fcos:
        preds = synth.make_fcos_preds(B, 800, 80, seed=1, device='cuda')
        crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
        dec = decode.FCOSDecoder(strides=synth.STRIDES)
    else:
        preds = synth.make_retina_preds(B, 800, 80, seed=1, device='cuda')
        crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
        dec = decode.RetinaDecoder(**synth.RETINA_KW)
    ann = synth.make_annotations(B, 100, 800, 80, seed=2).cuda()
    res = {}
    for what, fn in (('decode', lambda: dec(preds)), ('loss', lambda: crit(preds, ann))):
        with torch.no_grad():
            for _ in range(20):
                fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(200):
                fn()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / 200 * 1e3
            _lib.profile_start()
            for _ in range(50):
                fn()
            torch.cuda.synchronize()
            prof = _lib.profile_stop()
        res[what] = {'wall_ms': round(wall, 4),
                     'kernels_ms': {k: round(ms * n / 50, 4) for k, (n, ms) in prof.items()}}
    out[name] = res
print(json.dumps(out, indent=1))
