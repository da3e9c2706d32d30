import sys, torch
sys.path.insert(0, '.')
from b200det import synth, losses, _lib
dev = torch.device('cuda')
B = 64
preds = synth.make_retina_preds(B, 800, 80, seed=1, device=dev)
ann = synth.make_annotations(B, 100, 800, 80, seed=2).to(dev)
crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
req = [[t.detach().requires_grad_(True) for t in grp] for grp in preds]
def step():
    for grp in req:
        for t in grp: t.grad = None
    d = crit(req, ann)
    sum(d.values()).backward()
for _ in range(3): step()
torch.cuda.synchronize()
_lib.profile_start()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): step()
b.record(); torch.cuda.synchronize()
print('ms/step', a.elapsed_time(b) / 10, {k: (n, round(ms, 4)) for k, (n, ms) in _lib.profile_stop().items()})
