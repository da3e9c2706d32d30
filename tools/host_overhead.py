"""Host-side cost of one criterion / decoder call (small inputs so that the GPU is never the
bottleneck), with a cProfile breakdown.  python tools/host_overhead.py"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from b200det import synth, losses, decode  # noqa: E402

preds = synth.make_retina_preds(2, 128, 80, seed=0, device='cuda')
ann = synth.make_annotations(2, 10, 128, 80, seed=1).cuda()
crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
dec = decode.RetinaDecoder(**synth.RETINA_KW)


def loop(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return 1e6 * (t1 - t0) / n


with torch.no_grad():
    print('criterion host us/call', round(loop(lambda: crit(preds, ann)), 1))
    print('decoder   us/call (incl. its sync)', round(loop(lambda: dec(preds), 500), 1))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(2000):
        crit(preds, ann)
    pr.disable()
    torch.cuda.synchronize()
    pstats.Stats(pr).sort_stats('tottime').print_stats(14)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(500):
        dec(preds)
    pr.disable()
    pstats.Stats(pr).sort_stats('tottime').print_stats(14)
