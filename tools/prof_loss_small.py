"""No-grad RetinaLoss(GIoU) forward alone (BASELINE configs[1] shape) at small batches: wall time per call
with CUDA events.  python tools/prof_loss_small.py [--batches 16 8 4]"""
import argparse
import sys
import torch
sys.path.insert(0, '.')
from b200det import synth, losses
ap = argparse.ArgumentParser()
ap.add_argument('--batches', type=int, nargs='+', default=[16, 8, 4, 1])
ap.add_argument('--iters', type=int, default=200)
args = ap.parse_args()
crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
out = {}
for B in args.batches:
    preds = synth.make_retina_preds(B, 800, 80, seed=1, device='cuda')
    ann = synth.make_annotations(B, 100, 800, 80, seed=2).cuda()
    with torch.no_grad():
        for _ in range(10):
            crit(preds, ann)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.iters):
            crit(preds, ann)
        b.record()
        torch.cuda.synchronize()
    out[B] = round(a.elapsed_time(b) / args.iters, 4)
print(out)
