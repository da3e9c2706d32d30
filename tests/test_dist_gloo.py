"""world_size-2 gloo test (CPU) of the multi-GPU host logic: an image-sharded batch whose
{positives, loss sums} are all-reduced by the product's own helper reproduces the single-process
full-batch loss.  Per-rank sums come from the oracle here (no GPU); on the
B200 box the same helper all-reduces the sums written by b200det_loss_reduce over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from b200det import synth, losses
    from oracle import det_oracle as O
    torch.set_num_threads(2)
    B = 4
    preds = synth.make_retina_preds(B, 128, 8, seed=0)
    ann = synth.make_annotations(B, 12, 128, 8, seed=1, empty_images=(3,))
    per = B // world
    lo, hi = rank * per, (rank + 1) * per
    shard = [[t[lo:hi].contiguous() for t in grp] for grp in preds]
    with torch.no_grad():
        r = O.retina_loss(shard, ann[lo:hi], **synth.RETINA_KW, box_loss_type='GIoU')
    sums = torch.tensor([r['num_pos'], float(r['cls_sum']), float(r['reg_sum']), 0.],
                        dtype=torch.float64)
    losses._maybe_all_reduce(sums, True, None)          # the product's collective helper
    untouched = sums.clone()
    losses._maybe_all_reduce(untouched, False, None)    # sync_normalizer=False: no collective
    assert torch.equal(untouched, sums)
    np.save(os.path.join(out_dir, f'sums{rank}.npy'), sums.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_sums_reproduce_full_batch_loss(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    sys.path.insert(0, ROOT)
    from b200det import synth
    from oracle import det_oracle as O
    preds = synth.make_retina_preds(4, 128, 8, seed=0)
    ann = synth.make_annotations(4, 12, 128, 8, seed=1, empty_images=(3,))
    with torch.no_grad():
        full = O.retina_loss(preds, ann, **synth.RETINA_KW, box_loss_type='GIoU')
    s0 = np.load(tmp_path / 'sums0.npy')
    s1 = np.load(tmp_path / 'sums1.npy')
    assert np.array_equal(s0, s1)                       # every rank holds the global sums
    assert s0[0] == full['num_pos'] > 0
    # losses[i] = w_i * sum_i / positives  (b200det_loss_finish)
    np.testing.assert_allclose(s0[1] / s0[0], full['cls_loss'].item(), rtol=1e-6)
    np.testing.assert_allclose(s0[2] / s0[0], full['reg_loss'].item(), rtol=1e-6)


def test_bench_reference_arm_prints_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs) emits one JSON line."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--steps', '1', '--warmup', '0', '--ref-size', '128'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'images/s' and line['value'] > 0
    # the reference's own classes when baseline/fetch_ref.sh has vendored them, else the oracle port
    vendored = os.path.isfile(os.path.join(ROOT, 'baseline', '_ref', 'simpleAICV', 'detection',
                                           'losses.py'))
    assert line['cpu_baseline']['kind'] == ('reference' if vendored else 'port')
    assert line['e2e']['h2d_bytes_per_step'] == 0
    assert line['scaling'] == 'strong' and line['config']['global_batch'] == 256
