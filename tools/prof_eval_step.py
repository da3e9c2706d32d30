"""Per-kernel CUDA-event times of the fused evaluation step (b200det.fused.EvalStep) at the
benchmark shape.  python tools/prof_eval_step.py [--batch 256]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from b200det import synth, losses, decode, fused, _lib  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=256)
ap.add_argument('--iters', type=int, default=30)
args = ap.parse_args()
preds = synth.make_retina_preds(args.batch, 800, 80, seed=0, device='cuda')
ann = synth.make_annotations(args.batch, 100, 800, 80, seed=1).cuda()
step = fused.EvalStep(losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU'),
                      decode.RetinaDecoder(**synth.RETINA_KW))
for _ in range(5):
    step(preds, ann)
torch.cuda.synchronize()
_lib.profile_start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.iters):
    step(preds, ann)
e1.record()
torch.cuda.synchronize()
prof = _lib.profile_stop()
print('step ms', round(e0.elapsed_time(e1) / args.iters, 3),
      {k: round(v[1], 4) for k, v in prof.items()})
