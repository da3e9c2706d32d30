"""BASELINE configs[3] (FCOS, Objects365 shape: 365 classes, 1024x1024, batch 32, <= 200 GT) once per
stage: decode, no-grad loss, training forward + backward.  Run under ncu by tools/ncu_cfg4.sh."""
import sys
import torch
sys.path.insert(0, '.')
from b200det import synth, losses, decode, _lib
dev = torch.device('cuda')
B, S, C, G = 32, 1024, 365, 200
preds = synth.make_fcos_preds(B, S, C, seed=1, device=dev)
ann = synth.make_annotations(B, G, S, C, seed=2).to(dev)
crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
dec = decode.FCOSDecoder(strides=synth.STRIDES)
req = [[t.detach().requires_grad_(True) for t in grp] for grp in preds]


def step():
    for grp in req:
        for t in grp:
            t.grad = None
    dec(preds)
    with torch.no_grad():
        crit(preds, ann)
    sum(crit(req, ann).values()).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.profile_start()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    step()
b.record()
torch.cuda.synchronize()
print('ms/step', a.elapsed_time(b) / 10, {k: (n, round(ms, 4)) for k, (n, ms) in _lib.profile_stop().items()})
