"""BASELINE configs[3] (FCOS, Objects365 shape) evaluation step -- criterion(preds, annots) then
decoder(preds) -- with the sweep hand-over on and off: per-kernel CUDA-event times.
python tools/prof_cfg4_eval.py [--classes 365 --size 1024 --batch 32 --gt 200]"""
import argparse
import sys
import torch
sys.path.insert(0, '.')
from b200det import synth, losses, decode, _lib, _handoff
ap = argparse.ArgumentParser()
ap.add_argument('--classes', type=int, default=365)
ap.add_argument('--size', type=int, default=1024)
ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--gt', type=int, default=200)
ap.add_argument('--iters', type=int, default=30)
args = ap.parse_args()
dev = torch.device('cuda')
preds = synth.make_fcos_preds(args.batch, args.size, args.classes, seed=1, device=dev)
ann = synth.make_annotations(args.batch, args.gt, args.size, args.classes, seed=2).to(dev)
crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
dec = decode.FCOSDecoder(strides=synth.STRIDES)


def step():
    with torch.no_grad():
        d = crit(preds, ann)
    return d, dec(preds)


for on in (False, True):
    _handoff.ENABLED = on
    _handoff.reset()
    for _ in range(4):
        d, r = step()
    torch.cuda.synchronize()
    _lib.profile_start()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters):
        d, r = step()
    b.record()
    torch.cuda.synchronize()
    print('hand-over' if on else 'separate ', 'ms/step', round(a.elapsed_time(b) / args.iters, 4),
          {k: round(ms, 4) for k, (n, ms) in _lib.profile_stop().items()},
          {k: round(float(v), 6) for k, v in d.items()}, int((r[0] > 0).sum()))
