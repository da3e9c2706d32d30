"""Query-based decoders on one B200: b200det.decode.{DETRDecoder, DINODETRDecoder} against the
reference-style path (torch activation on the GPU, D2H of every tensor, NumPy arg-max / sort / NMS
on the host = oracle.query_decode, the restatement of decode.py:367-594).  CUDA events + wall clock;
the calls end with the result on the host, as the reference's do.

    python tests/perf_queries.py [--batch 16]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
from b200det import decode  # noqa: E402
from oracle import det_oracle as O  # noqa: E402  (the reference-style leg; lives under tests/ for that reason)


def wall(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=16)
    ap.add_argument('--iters', type=int, default=50)
    args = ap.parse_args()
    B = args.batch
    gen = torch.Generator(device='cuda').manual_seed(0)
    sizes = [[800, 1333]] * B
    out = {'batch': B}
    for name, Q, C in (('DETR', 100, 81), ('DINO-DETR', 900, 80)):
        cls = torch.randn((6, B, Q, C), generator=gen, device='cuda') * 2
        reg = torch.rand((6, B, Q, 4), generator=gen, device='cuda') * 0.5
        if name == 'DETR':
            dec = decode.DETRDecoder(num_classes=C - 1, nms_type='python_nms')
            ours = lambda: dec([cls, reg], sizes)   # noqa: E731
            ref = lambda: O.query_decode(cls[-1], reg[-1], sizes, 'softmax', num_classes=C - 1,   # noqa: E731
                                         nms_type='python_nms')
        else:
            dec = decode.DINODETRDecoder()
            ours = lambda: dec({'pred_logits': cls[-1], 'pred_boxes': reg[-1]}, sizes)   # noqa: E731
            ref = lambda: O.query_decode(cls[-1], reg[-1], sizes, 'sigmoid', topn=300,   # noqa: E731
                                         nms_type='python_nms')
        ms_ours = wall(ours, args.iters)
        ms_ref = wall(ref, max(3, args.iters // 10))
        out[name] = {'queries': Q, 'channels': C, 'b200det_ms': round(ms_ours, 4),
                     'torch_numpy_ms': round(ms_ref, 3),
                     'images_per_s': round(B / ms_ours * 1e3, 1),
                     'speedup': round(ms_ref / ms_ours, 1)}
    print(json.dumps(out))


if __name__ == '__main__':
    main()
