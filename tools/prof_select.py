"""Where one decoder call spends its time at small batches (BASELINE configs 0-3 are latency-bound).

    python tools/prof_select.py [--reps 200]

Per config: wall time of decoder(preds) (host call to host result), CUDA-event time of the two
kernels, and the select kernel's phase breakdown from %globaltimer stamps written by the leader CTA
of every image (b200det_select_stamps; median over images and repetitions).  A/B knobs are read by
the library once per process: B200DET_SELECT_SLICES=1|2|4|8, B200DET_SELECT_BITONIC=1.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from b200det import synth, decode, _lib  # noqa: E402

PHASES = ['histograms', 'candidate_list', 'order', 'box_decode', 'nms', 'outputs']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--reps', type=int, default=200)
    lib = _lib.load()
    out = {'knobs': {k: os.environ.get(k) for k in ('B200DET_SELECT_SLICES', 'B200DET_SELECT_BITONIC')}}
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    for name, B, fcos, S, C in (('retina_b1', 1, False, 800, 80), ('retina_b16', 16, False, 800, 80),
                                ('retina_b32', 32, False, 800, 80), ('retina_b64', 64, False, 800, 80),
                                ('retina_b128', 128, False, 800, 80), ('retina_b256', 256, False, 800, 80),
                                ('fcos_b16', 16, True, 800, 80), ('fcos_1024_c365_b32', 32, True, 1024, 365)):
        if args.only and name not in args.only.split(','):
            continue
        if fcos:
            preds = synth.make_fcos_preds(B, S, C, seed=1, device='cuda')
            dec = decode.FCOSDecoder(strides=synth.STRIDES)
        else:
            preds = synth.make_retina_preds(B, S, C, seed=1, device='cuda')
            dec = decode.RetinaDecoder(**synth.RETINA_KW)
        reps = args.reps if B < 128 else max(10, args.reps // 10)
        for _ in range(10):
            dec(preds)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            dec(preds)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / reps * 1e3
        _lib.profile_start()
        for _ in range(reps):
            dec(preds)
        torch.cuda.synchronize()
        prof = _lib.profile_stop()
        stamps = torch.zeros((B, 16), dtype=torch.int64, device='cuda')
        lib.b200det_select_stamps(stamps.data_ptr())
        rows = []
        for _ in range(min(reps, 50)):
            dec(preds)
            torch.cuda.synchronize()
            rows.append(stamps.cpu().numpy()[:, :11].copy())
        lib.b200det_select_stamps(None)
        st = np.stack(rows).astype(np.float64)          # [reps, B, 11] ns
        d = np.diff(st[:, :, :7], axis=2) / 1e3         # us per phase

        def span(a, b):
            return round(float(np.median(st[:, :, b] - st[:, :, a])) / 1e3, 2)
        out[name] = {
            'wall_ms': round(wall, 4),
            'kernels_ms': {k: round(ms, 4) for k, (n, ms) in prof.items()},
            'select_phases_us': {p: round(float(np.median(d[:, :, i])), 2) for i, p in enumerate(PHASES)},
            'front_end_detail_us': {'prologue_until_predecessor_done': span(0, 10),
                                    'key_pass_1_histogram': span(10, 1), 'dsmem_histogram_sum': span(1, 7),
                                    'find_cut': span(7, 8), 'key_pass_2_collect': span(8, 9),
                                    'append_to_leader_and_cluster_sync': span(9, 2)},
            'select_leader_total_us': round(float(np.median(st[:, :, 6] - st[:, :, 0])) / 1e3, 2),
            'select_all_images_span_us': round(float(np.median(st[:, :, 6].max(1) - st[:, :, 0].min(1))) / 1e3, 2),
        }
        del preds
        torch.cuda.empty_cache()
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
