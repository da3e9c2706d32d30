"""One decoder call per small-batch config (after warm-up): run under
`ncu --metrics gpu__time_duration.sum` to see the select chain's kernels one by one."""
import sys
import torch
sys.path.insert(0, '.')
from b200det import synth, decode
for name, B, fcos in (('retina_b1', 1, False), ('retina_b16', 16, False), ('fcos_b16', 16, True)):
    if fcos:
        preds = synth.make_fcos_preds(B, 800, 80, seed=1, device='cuda')
        dec = decode.FCOSDecoder(strides=synth.STRIDES)
    else:
        preds = synth.make_retina_preds(B, 800, 80, seed=1, device='cuda')
        dec = decode.RetinaDecoder(**synth.RETINA_KW)
    for _ in range(3):
        dec(preds)
    torch.cuda.synchronize()
