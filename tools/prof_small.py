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
for name, B, fcos, S, C, G in (('retina_b1', 1, False, 800, 80, 100),
                               ('retina_b16', 16, False, 800, 80, 100),
                               ('fcos_b16', 16, True, 800, 80, 100),
                               ('fcos_1024_c365_b32', 32, True, 1024, 365, 200)):
    if fcos:
        preds = synth.make_fcos_preds(B, S, C, seed=1, device='cuda')
        crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
        dec = decode.FCOSDecoder(strides=synth.STRIDES)
    else:
        preds = synth.make_retina_preds(B, S, C, seed=1, device='cuda')
        crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
        dec = decode.RetinaDecoder(**synth.RETINA_KW)
    ann = synth.make_annotations(B, G, S, C, seed=2).cuda()
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
    # the same calls replayed from CUDA graphs (no Python, no launch gaps): no-grad loss forward and
    # the training forward + backward (tests/test_gpu_graphs.py checks the replays bit for bit)
    req = [[t.clone().requires_grad_(True) for t in grp] for grp in preds]

    def fwd():
        with torch.no_grad():
            return crit(preds, ann)

    def train():
        sum(crit(req, ann).values()).backward()

    def cleared():
        for grp in req:
            for t in grp:
                t.grad = None
        train()

    for what, fn in (('loss', fwd), ('loss_fwd_bwd', train)):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for grp in req:
            for t in grp:
                t.grad = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            fn()
        row = res.setdefault(what, {})
        for key, call in (('eager_ms', fwd if what == 'loss' else cleared), ('graph_replay_ms', graph.replay)):
            for _ in range(10):
                call()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(100):
                call()
            b.record()
            torch.cuda.synchronize()
            row[key] = round(a.elapsed_time(b) / 100, 4)
        del graph
    out[name] = res
print(json.dumps(out, indent=1))
