"""Host-side profile of one eager training step of the criterion at a small batch (it is host-bound
there): python tools/prof_train_host.py"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from b200det import synth, losses  # noqa: E402

for name, fcos in (('retina_b16', False), ('fcos_b16', True)):
    if fcos:
        preds = synth.make_fcos_preds(16, 800, 80, seed=1, device='cuda')
        crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
    else:
        preds = synth.make_retina_preds(16, 800, 80, seed=1, device='cuda')
        crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
    ann = synth.make_annotations(16, 100, 800, 80, seed=2).cuda()
    req = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]

    def step():
        for grp in req:
            for t in grp:
                t.grad = None
        sum(crit(req, ann).values()).backward()

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200):
        step()
    torch.cuda.synchronize()
    print(name, 'ms/step', round((time.perf_counter() - t0) / 200 * 1e3, 4))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(200):
        step()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats('tottime').print_stats(16)
