#!/usr/bin/env python
"""bench.py -- det loss+decode+NMS images/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200det|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic RetinaNet-R50 head outputs that
are already resident in HBM: RetinaLoss (IoU assignment + focal + GIoU box loss, forward) followed
by RetinaDecoder (score threshold, top-1000, box decode, NMS, 100 detections), exactly the two calls
the reference's eval loop makes per batch (tools/scripts.py:733-740).  Workload = BASELINE.json
configs[4] ("RetinaNet-R50 loss+decode sharded by image at batch 256"): COCO 80 classes, 800x800,
9 anchors/location, <= 100 GT boxes per image, 256 images PER GPU (weak scaling: every rank holds
its own shard; the only collective is the NCCL all-reduce of {positives, loss sums}).

Prints ONE JSON line (rank 0).  `value` is kernel-path throughput with inputs in HBM; `e2e` is the
same step through the public classes starting from pinned HOST buffers (H2D of every head output
inside the timed region, D2H of losses and detections); `roofline` is the dominant kernel against
the measured HBM copy peak (MEASURED_PEAKS.json); `cpu_baseline` is the oracle port of the
reference's torch/NumPy CPU path timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SIZE, NUM_CLASSES, MAX_GT = 800, 80, 100
METRIC = 'det loss+decode+NMS images/s'
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def rows_per_image():
    from b200det import synth
    return sum(p * p * 9 for p in synth.pyramid_sizes(SIZE))


def algorithmic_bytes_per_image():
    """loss_fwd = 4NC + 16N + 20G ; decode = 4NC + 16N + 2400."""
    n, c, g = rows_per_image(), NUM_CLASSES, MAX_GT
    loss = 4 * n * c + 16 * n + 20 * g
    dec = 4 * n * c + 16 * n + 2400
    return loss, dec


def hbm_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return FALLBACK_HBM_GBS, 'fallback (B200_PROFILING.md)'


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            'hw_slowdown': getattr(nv, 'nvmlClocksEventReasonHwSlowdown', 0x8),
            'hw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown', 0x40),
            'sw_thermal_slowdown': getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown', 0x20),
            'sw_power_cap': getattr(nv, 'nvmlClocksEventReasonSwPowerCap', 0x4),
            'hw_power_brake': getattr(nv, 'nvmlClocksEventReasonHwPowerBrakeSlowdown', 0x80),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {
            'sm_mhz': s[len(s) // 2] if s else None,
            'sm_max_mhz': self.max_mhz,
            'reasons': sorted(self.reasons),
            'samples': len(s),
        }


# --------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's torch + NumPy CPU path
# --------------------------------------------------------------------------------------------
def cpu_step(preds, ann):
    """One reference-style eval step on the host: criterion(outs, annots); decoder(outs)."""
    import torch
    from b200det import synth
    from oracle import det_oracle as O
    with torch.no_grad():
        out = O.retina_loss(preds, ann, **synth.RETINA_KW, box_loss_type='GIoU')
        res, _ = O.retina_decode(preds, **synth.RETINA_KW)
    return out, res


def time_cpu(images_per_step, steps, warmup, seed=0, size=SIZE):
    import torch
    from b200det import synth
    torch.set_num_threads(os.cpu_count() or 1)
    preds = synth.make_retina_preds(images_per_step, size, NUM_CLASSES, seed=seed)
    ann = synth.make_annotations(images_per_step, MAX_GT, size, NUM_CLASSES, seed=seed + 1)
    for _ in range(warmup):
        cpu_step(preds, ann)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(preds, ann)
    dt = time.perf_counter() - t0
    return images_per_step * steps / dt, dt, torch.get_num_threads()


def run_reference(args, out):
    """--impl reference: the reference's own CPU implementation of the path (oracle port: the
    reference is pure Python and cannot travel to the GPU box), all host threads, bounded sample
    per step.  Under torchrun only rank 0 works."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    per_step = 2
    value, dt, threads = time_cpu(per_step, args.steps, min(args.warmup, 1), size=args.ref_size)
    loss_b, dec_b = algorithmic_bytes_per_image()
    line = {
        'impl': 'reference',
        'metric': METRIC,
        'value': value,
        'unit': 'images/s',
        'n_gpus': args.gpus,
        'steps': args.steps,
        'warmup': args.warmup,
        'ms_per_step': 1e3 * dt / max(args.steps, 1),
        'higher_is_better': True,
        'scaling': 'weak',
        'vs_baseline': None,
        'dtype': 'f32',
        'data': 'synthetic',
        # the arm's own config (what the sample is a sample OF); the sample size is below
        'config': workload_config(args.batch, args.gpus),
        'sample_images_per_step': per_step,
        'cpu_baseline': {
            'value': value,
            'unit': 'images/s',
            'cores': threads,
            'kind': 'port',
            'sample': f'{per_step} images per step x {args.steps} steps of the same workload; '
                      'torch-CPU loss uses all threads, NumPy decode/NMS is single-threaded',
        },
        'e2e': {'value': value, 'unit': 'images/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    out.emit(json.dumps(line))


def workload_config(batch_per_gpu, n_gpus, exchange='nccl'):
    loss_b, dec_b = algorithmic_bytes_per_image()
    how = ('4 doubles per step exchanged over NVLink peer memory inside the reduction kernel'
           if exchange == 'p2p' else 'NCCL all-reduce of 4 doubles per step')
    if n_gpus == 1:
        how = 'no exchange'

    return {
        'workload': 'BASELINE configs[4]: RetinaNet-R50 head outputs, RetinaLoss(GIoU) forward + '
                    'RetinaDecoder(python_nms), COCO 80 cls, 800x800, 9 anchors/loc, <=100 GT/img',
        'images_per_gpu': batch_per_gpu,
        'global_batch': batch_per_gpu * n_gpus,
        'rows_per_image': rows_per_image(),
        'algorithmic_bytes_per_image': loss_b + dec_b,
        'parallelism': f'image-sharded x{n_gpus}, {how}',
        'l2_policy': 'inputs (>= 10 GB per GPU at the default batch) far exceed the 126 MB L2',
    }


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def run_b200(args, out):
    import torch
    import torch.distributed as dist
    import b200det
    from b200det import synth, losses, decode, _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device; there is no CPU path (use --impl reference '
                           'for the host baseline)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    _lib.load()

    B = args.batch
    preds = synth.make_retina_preds(B, SIZE, NUM_CLASSES, seed=100 + rank, device=dev)
    ann = synth.make_annotations(B, MAX_GT, SIZE, NUM_CLASSES, seed=200 + rank).to(dev)
    # the normaliser's exchange: fused into the reduction kernel over NVLink peer memory when every
    # rank can map every peer (decided collectively), else torch.distributed / NCCL
    exchange = 'nccl'
    crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU',
                             sync_normalizer=distributed)
    if distributed and args.sync in ('auto', 'p2p'):
        ok = torch.ones(1, device=dev)
        try:
            from b200det.peer import PeerExchange
            crit._peer = PeerExchange(None, dev)
        except Exception as exc:   # noqa: BLE001 -- any failure means "use NCCL"
            print(f'[bench] rank {rank}: peer exchange unavailable ({exc}); using NCCL',
                  file=sys.stderr)
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() > 0:
            crit.sync_normalizer = 'p2p'
            exchange = 'p2p'
        elif args.sync == 'p2p':
            raise RuntimeError('--sync p2p requested but the peer exchange could not be set up')
    dec = decode.RetinaDecoder(**synth.RETINA_KW)

    def step():
        with torch.no_grad():
            d = crit(preds, ann)
            r = dec(preds)
        return d, r

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    _lib.profile_start()
    launches0 = _lib.launch_count()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for i in range(args.steps):
        # per-kernel CUDA events on every `profile_every`-th step of the timed region (bracketing
        # every launch costs ~14 event records per step)
        if args.profile_every > 1:
            (_lib.profile_resume if i % args.profile_every == 0 else _lib.profile_pause)()
        d, r = step()
    stop.record()
    barrier()
    elapsed_ms = start.elapsed_time(stop)
    launches = _lib.launch_count() - launches0
    kernels = _lib.profile_stop()
    clocks = sampler.stop()

    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = B * world * args.steps / (elapsed_ms / 1e3)

    # ---- optional extension: one sweep over cls for loss + decode (b200det.fused.EvalStep) ----
    fused_info = None
    if not args.no_fused:
        from b200det import fused
        fstep = fused.EvalStep(crit, dec)
        for _ in range(3):
            fstep(preds, ann)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            fstep(preds, ann)
        f1.record()
        barrier()
        ft = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(ft, op=dist.ReduceOp.MAX)
        fms = float(ft.item()) / args.steps
        fused_info = {
            'value': B * world / (fms / 1e3),
            'unit': 'images/s',
            'ms_per_step': fms,
            'note': 'NOT the drop-in call structure: fused.EvalStep(criterion, decoder) reads cls '
                    'once for loss + decode (one call instead of the reference\'s two)',
        }

    # ---- e2e: host buffers in, host results out -------------------------------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, preds, ann, crit, dec, dev, distributed, world)

    # ---- roofline of the dominant kernel ----------------------------------------------------
    peak, peak_src = hbm_peak()
    n, c = rows_per_image(), NUM_CLASSES
    alg = {
        'focal_loss': B * 4 * n * c,
        'score_argmax': B * 4 * n * c,
        'assign': B * (16 * n + 20 * MAX_GT),
        'select_decode_nms': B * (8 * n + 16 * 1000 + 2400),
    }
    dom = max((k for k in kernels if k in ('focal_loss', 'score_argmax')),
              key=lambda k: kernels[k][1])
    dom_ms = kernels[dom][1]
    achieved = alg[dom] / (dom_ms / 1e3) / 1e9
    loss_b, dec_b = algorithmic_bytes_per_image()
    step_gbs = B * (loss_b + dec_b) / (ms_per_step / 1e3) / 1e9
    traffic = ncu_traffic()

    line = {
        'metric': METRIC,
        'value': value,
        'unit': 'images/s',
        'n_gpus': world,
        'steps': args.steps,
        'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step,
        'higher_is_better': True,
        'scaling': 'weak',
        'vs_baseline': None,
        'dtype': 'f32',
        'data': 'synthetic',
        'config': workload_config(B, world, exchange),
        'roofline': {
            'bound': 'hbm',
            'kernel': dom,
            'achieved': achieved,
            'peak': peak,
            'unit': 'GB/s',
            'frac': achieved / peak,
            'traffic': traffic.get(dom),
            'peak_source': peak_src,
            'algorithmic_bytes_per_launch': alg[dom],
            'kernel_ms': dom_ms,
        },
        'step_roofline': {
            'algorithmic_GBps': step_gbs,
            'frac_of_hbm_peak': step_gbs / peak,
            'note': 'whole step (loss fwd + decode + NMS, 80.70 MB/image) against the HBM peak',
        },
        'kernels_ms': {k: round(v[1], 4) for k, v in kernels.items()},
        'clocks': clocks,
        'gpu_launches': launches,
        'e2e': e2e,
        'fused_eval_step': fused_info,
        'loss': {k: float(v.item()) for k, v in d.items()},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, threads = time_cpu(2, args.cpu_steps, 1)
        line['cpu_baseline'] = {
            'value': v,
            'unit': 'images/s',
            'cores': threads,
            'kind': 'port',
            'sample': f'2 images per step x {args.cpu_steps} steps of the same workload '
                      f'({dt:.1f} s); torch-CPU loss on all threads, NumPy decode single-threaded',
        }
    if rank == 0:
        out.emit(json.dumps(line))
    if distributed:
        dist.destroy_process_group()


def run_e2e(args, preds, ann, crit, dec, dev, distributed, world):
    """Same step through the public classes, but starting from pinned host buffers every step."""
    import torch
    import torch.distributed as dist
    B = args.e2e_batch or args.batch
    try:
        import psutil
        need = B * algorithmic_bytes_per_image()[0] * 1.1
        if psutil.virtual_memory().available < 3 * need * max(world, 1):
            B = max(8, B // 8)
    except Exception:
        pass
    # allocate the pinned staging buffers from the NUMA node next to this GPU (first touch), then
    # give the thread its old CPU mask back (the CPU baseline must keep all host cores)
    old_mask = os.sched_getaffinity(0)
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(dev.index or 0))
        print(f'[bench] rank {os.environ.get("RANK", "0")}: pinned buffers allocated from CPUs '
              f'{sorted(os.sched_getaffinity(0))[:4]}.. ({len(os.sched_getaffinity(0))} cpus)',
              file=sys.stderr)
    except Exception as e:  # restricted cpuset, no NVML: keep the default placement
        print(f'[bench] NUMA binding skipped: {e}', file=sys.stderr)
    try:
        host = [[t[:B].cpu().pin_memory() for t in grp] for grp in preds]
        host_ann = ann[:B].cpu().pin_memory()
    finally:
        os.sched_setaffinity(0, old_mask)
    h2d = sum(t.numel() * t.element_size() for grp in host for t in grp)
    h2d += host_ann.numel() * host_ann.element_size()
    d2h = 6 * B * 100 * 4 + 2 * 4

    def step():
        with torch.no_grad():
            p = [[t.to(dev, non_blocking=True) for t in grp] for grp in host]
            a = host_ann.to(dev, non_blocking=True)
            d = crit(p, a)
            r = dec(p)
            vals = torch.stack([d['cls_loss'], d['reg_loss']]).cpu()   # D2H of the losses
        return vals, r

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        step()
    barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    return {
        'value': B * world * args.e2e_steps / dt,
        'unit': 'images/s',
        'h2d_bytes_per_step': h2d,
        'd2h_bytes_per_step': d2h,
        'images_per_gpu_per_step': B,
        'steps': args.e2e_steps,
        'note': 'pinned host -> HBM copy of all head outputs + annotations, loss, decode+NMS, D2H '
                'of losses and detections, every step; PCIe-bound',
    }


class StdoutToStderr:
    """Everything a library prints to fd 1 during the run (NCCL's version banner, ...) goes to
    stderr, so that stdout carries exactly ONE line: the JSON result."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + '\n').encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200det', choices=['b200det', 'reference'])
    ap.add_argument('--batch', type=int, default=256, help='images per GPU per step')
    ap.add_argument('--e2e-batch', type=int, default=0)
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--cpu-steps', type=int, default=40)
    ap.add_argument('--ref-size', type=int, default=SIZE,
                    help='image size of the --impl reference sample (tests use a small one)')
    ap.add_argument('--sync', default='auto', choices=['auto', 'p2p', 'nccl'],
                    help='N > 1: how the loss normaliser crosses GPUs')
    ap.add_argument('--no-fused', action='store_true')
    ap.add_argument('--profile-every', type=int, default=1)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    with StdoutToStderr() as out:
        if args.impl == 'reference':
            run_reference(args, out)
        else:
            run_b200(args, out)


if __name__ == '__main__':
    main()
