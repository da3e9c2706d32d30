#!/usr/bin/env python
"""bench.py -- det loss+decode+NMS images/s on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200det|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic RetinaNet-R50 head outputs that
are already resident in HBM: RetinaLoss (IoU assignment + focal + GIoU box loss, forward) followed
by RetinaDecoder (score threshold, top-1000, box decode, NMS, 100 detections), exactly the two calls
the reference's eval loop makes per batch (tools/scripts.py:733-740).  Workload = BASELINE.json
configs[4] ("RetinaNet-R50 loss+decode sharded by image at batch 256 across 1/2/4/8 B200"): COCO 80
classes, 800x800, 9 anchors/location, <= 100 GT boxes per image, ONE global batch of 256 images
sharded by image: rank r holds images [r*256/N, (r+1)*256/N)  (STRONG scaling -- the
global batch is the same at every N, built from per-image seeds so that it is the same batch bit for
bit however it is sharded).  The only collective is the exchange of {positives, loss sums} (4
doubles): the loss every rank returns is the reference's single-process loss of the unsharded batch
(losses.py:231-259).  At N > 1 the line also carries `weak_scaling` (256 images PER GPU, round 1's
definition) and `parity_check` (sharded vs unsharded loss and labels on a 4*N-image batch; the run
fails if they disagree).

Prints ONE JSON line (rank 0).  `value` is kernel-path throughput with inputs in HBM; `e2e` is the
same step through the public classes starting from pinned HOST buffers (H2D of every head output
inside the timed region, D2H of losses and detections); `roofline` is the dominant kernel against
the measured HBM copy peak (MEASURED_PEAKS.json); `cpu_baseline` is the reference's own classes
(baseline/_ref, vendored unmodified by baseline/fetch_ref.sh; the oracle port if absent) timed on
this box's host cores on a bounded sample; `configs` (N = 1) times BASELINE configs[0..3] and the
training step of configs[4].
"""
import argparse
import json
import os
import gc
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SIZE, NUM_CLASSES, MAX_GT = 800, 80, 100
GLOBAL_BATCH = 256
METRIC = 'det loss+decode+NMS images/s'
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def rows_per_image(size=SIZE, per_loc=9):
    from b200det import synth
    return sum(p * p * per_loc for p in synth.pyramid_sizes(size))


def stage_bytes(n, c, g, k):
    """SURVEY.md section 8(d), per image: loss_fwd = 4NC + 16N + 4Nk + 20G ; loss_bwd (extra) =
    4NC + 16N + 4Nk ; decode = 4NC + 16N + 4Nk + 2400  (k = 1 with a centre-ness head)."""
    loss = 4 * n * c + 16 * n + 4 * n * k + 20 * g
    bwd = 4 * n * c + 16 * n + 4 * n * k
    dec = 4 * n * c + 16 * n + 4 * n * k + 2400
    return loss, bwd, dec


def algorithmic_bytes_per_image():
    loss, _, dec = stage_bytes(rows_per_image(), NUM_CLASSES, MAX_GT, 0)
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


_PROBE_SRC = r"""
import json, sys, time
import pynvml as nv
nv.nvmlInit()
h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
period = float(sys.argv[2])
names = {'hw_slowdown': 0x8, 'hw_thermal_slowdown': 0x40, 'sw_thermal_slowdown': 0x20,
         'sw_power_cap': 0x4, 'hw_power_brake': 0x80}
samples, reasons = [], set()
max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
sys.stdout.write('ready\n'); sys.stdout.flush()
import select
while True:
    samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
    try:
        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
    except Exception:
        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
    for k, bit in names.items():
        if mask & bit:
            reasons.add(k)
    if select.select([sys.stdin], [], [], period)[0]:
        break
s = sorted(samples)
print(json.dumps({'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': max_mhz,
                  'reasons': sorted(reasons), 'samples': len(s)}))
"""


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU with NVML while the timed region runs -- in a
    separate PROCESS: a sampling thread inside the benchmark process takes the GIL every few
    milliseconds, and with one process per GPU every such hiccup stalls ALL ranks at the next
    exchange of the loss normaliser (measured at 8 GPUs, 32 images per rank)."""

    def __init__(self, index, period=0.005):
        import subprocess
        self.proc = None
        try:
            # NVML indexes physical GPUs: honour CUDA_VISIBLE_DEVICES
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            if vis:
                ids = [v.strip() for v in vis.split(',') if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    index = int(ids[index])
            self.proc = subprocess.Popen([sys.executable, '-c', _PROBE_SRC, str(index), str(period)],
                                         stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def start(self):
        """Blocks until the probe has initialised NVML and taken up sampling."""
        if self.proc is not None:
            try:
                if self.proc.stdout.readline().strip() != 'ready':
                    self.proc.kill()
                    self.proc = None
            except Exception:
                self.proc = None

    def stop(self):
        none = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return none
        try:
            out, _ = self.proc.communicate('stop\n', timeout=10)
            return json.loads(out.strip().splitlines()[-1])
        except Exception:
            try:
                self.proc.kill()
            except Exception:
                pass
            return none


# --------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own classes on the host (baseline/_ref), or the
# oracle port of its torch + NumPy CPU path when the vendored files are absent
# --------------------------------------------------------------------------------------------
def load_reference():
    """(kind, losses_module, decode_module): the UNMODIFIED reference vendored by
    baseline/fetch_ref.sh (kind 'reference'), else None -> the oracle port (kind 'port')."""
    try:
        from baseline import refarm
        if refarm.available():
            ref_losses, ref_decode = refarm.load()
            return 'reference', ref_losses, ref_decode
    except Exception as exc:   # noqa: BLE001
        print(f'[bench] vendored reference unusable ({exc}); timing the oracle port', file=sys.stderr)
    return 'port', None, None


def make_cpu_step(kind_cfg, ref):
    """Returns step(preds, ann) -> None running criterion(outs, annots); decoder(outs) like
    tools/scripts.py:733-740 on CPU tensors.  kind_cfg: 'retina', 'retina_decode', 'retina_loss',
    'fcos'."""
    import torch
    from b200det import synth
    kind, ref_losses, ref_decode = ref
    want_loss = kind_cfg != 'retina_decode'
    want_dec = kind_cfg != 'retina_loss'
    if kind == 'reference':
        if kind_cfg == 'fcos':
            crit = ref_losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
            dec = ref_decode.FCOSDecoder(strides=synth.STRIDES)
        else:
            crit = ref_losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
            dec = ref_decode.RetinaDecoder(**synth.RETINA_KW)

        def step(preds, ann):
            with torch.no_grad():
                if want_loss:
                    crit(preds, ann)
                if want_dec:
                    dec(preds)
        return step

    from oracle import det_oracle as O

    def step(preds, ann):
        with torch.no_grad():
            if kind_cfg == 'fcos':
                if want_loss:
                    O.fcos_loss(preds, ann, synth.STRIDES, synth.MI)
                if want_dec:
                    O.fcos_decode(preds, synth.STRIDES)
            else:
                if want_loss:
                    O.retina_loss(preds, ann, **synth.RETINA_KW, box_loss_type='GIoU')
                if want_dec:
                    O.retina_decode(preds, **synth.RETINA_KW)
    return step


def time_cpu(ref, kind_cfg, images_per_step, steps, warmup, size=SIZE, classes=NUM_CLASSES,
             max_gt=MAX_GT, seed=0):
    import torch
    from b200det import synth
    torch.set_num_threads(os.cpu_count() or 1)
    if kind_cfg == 'fcos':
        preds = synth.make_fcos_preds(images_per_step, size, classes, seed=seed)
    else:
        preds = synth.make_retina_preds(images_per_step, size, classes, seed=seed)
    ann = synth.make_annotations(images_per_step, max_gt, size, classes, seed=seed + 1)
    step = make_cpu_step(kind_cfg, ref)
    for _ in range(warmup):
        step(preds, ann)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(preds, ann)
    dt = time.perf_counter() - t0
    return images_per_step * steps / dt, dt, torch.get_num_threads()


def cpu_sample_note(kind, per_step, steps, dt):
    what = ("the reference's own RetinaLoss + RetinaDecoder (baseline/_ref, unmodified)"
            if kind == 'reference' else 'oracle port of the reference path')
    return (f'{per_step} images per step x {steps} steps of the same workload ({dt:.1f} s), {what}; '
            'torch-CPU loss on all threads, NumPy decode/NMS single-threaded')


def run_reference(args, out):
    """--impl reference: the reference's own CPU implementation of the path, all host threads,
    bounded sample per step.  Under torchrun only rank 0 works."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    ref = load_reference()
    per_step = 2
    value, dt, threads = time_cpu(ref, 'retina', per_step, args.steps, min(args.warmup, 1),
                                  size=args.ref_size)
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
        'scaling': args.scaling,
        'vs_baseline': None,
        'dtype': 'f32',
        'data': 'synthetic',
        # the arm's own config (what the sample is a sample OF); the sample size is below
        'config': workload_config(local_batch(args, args.gpus), args.gpus, args.scaling),
        'sample_images_per_step': per_step,
        'cpu_baseline': {
            'value': value,
            'unit': 'images/s',
            'cores': threads,
            'kind': ref[0],
            'sample': cpu_sample_note(ref[0], per_step, args.steps, dt),
        },
        'e2e': {'value': value, 'unit': 'images/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    out.emit(json.dumps(line))


def local_batch(args, world):
    if args.batch:
        return args.batch
    if args.scaling == 'weak':
        return GLOBAL_BATCH
    if GLOBAL_BATCH % world:
        raise SystemExit(f'--gpus {world} does not divide the global batch of {GLOBAL_BATCH}')
    return GLOBAL_BATCH // world


def workload_config(batch_per_gpu, n_gpus, scaling):
    loss_b, dec_b = algorithmic_bytes_per_image()
    return {
        'workload': 'BASELINE configs[4]: RetinaNet-R50 head outputs, RetinaLoss(GIoU) forward + '
                    'RetinaDecoder(python_nms), COCO 80 cls, 800x800, 9 anchors/loc, <=100 GT/img',
        'images_per_gpu': batch_per_gpu,
        'global_batch': batch_per_gpu * n_gpus,
        'rows_per_image': rows_per_image(),
        'algorithmic_bytes_per_image': loss_b + dec_b,
        'parallelism': f'one global batch sharded by image x{n_gpus}; loss normalised by the '
                       'global positive count (4 doubles exchanged per step)'
                       if scaling == 'strong' else f'{batch_per_gpu} images per GPU x{n_gpus}',
        'l2_policy': 'per-GPU inputs (>= 1.29 GB at 32 images) far exceed the 126 MB L2',
        'call_structure': 'the reference\'s two calls per step, criterion(preds, annots) then '
                          'decoder(preds), on the same tensors; the criterion\'s sweep also writes the '
                          'decoder\'s keys (computed every step, verified on the device; b200det/_handoff.py)',
    }


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
class Runner:
    """Process-group plumbing + the timed loop shared by the headline, weak-scaling and fused runs."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from b200det import _lib
        self.torch, self.dist, self.lib = torch, dist, _lib
        self.args = args
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.rank = int(os.environ.get('RANK', '0'))
        self.local_rank = int(os.environ.get('LOCAL_RANK', '0'))
        if not torch.cuda.is_available():
            raise RuntimeError('bench.py needs a CUDA device; there is no CPU path (use --impl '
                               'reference for the host baseline)')
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device('cuda', self.local_rank)
        self.distributed = self.world > 1
        if self.distributed:
            os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
            dist.init_process_group('nccl', device_id=self.dev)
        _lib.load()
        self.exchange = 'none'
        self.cpus = None
        if self.distributed and not args.no_pin:
            # one disjoint slice of the allowed CPUs per rank: the launcher starts all ranks on the
            # same cpuset, and a migrating / preempted main thread on one rank delays all of them
            try:
                allowed = sorted(os.sched_getaffinity(0))
                per = max(1, len(allowed) // self.world)
                mine = allowed[self.local_rank * per:(self.local_rank + 1) * per]
                if mine:
                    os.sched_setaffinity(0, mine)
                    self.cpus = len(mine)
            except Exception as exc:   # noqa: BLE001
                print(f'[bench] CPU pinning skipped: {exc}', file=sys.stderr)

    def barrier(self):
        if self.distributed:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.distributed:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def make_criterion(self, sync=True):
        """RetinaLoss(GIoU) whose normaliser crosses the ranks: inside the reduction kernel over
        NVLink peer memory when every rank can map every peer (decided collectively), else NCCL."""
        from b200det import losses, synth
        torch, dist, args = self.torch, self.dist, self.args
        crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU',
                                 sync_normalizer=self.distributed and sync)
        if not (self.distributed and sync):
            return crit
        self.exchange = 'nccl'
        if args.sync in ('auto', 'p2p'):
            ok = torch.ones(1, device=self.dev)
            try:
                from b200det.peer import PeerExchange
                crit._peer = PeerExchange(None, self.dev)
            except Exception as exc:   # noqa: BLE001 -- any failure means "use NCCL"
                print(f'[bench] rank {self.rank}: peer exchange unavailable ({exc}); using NCCL',
                      file=sys.stderr)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() > 0:
                crit.sync_normalizer = 'p2p'
                self.exchange = 'p2p'
            elif args.sync == 'p2p':
                raise RuntimeError('--sync p2p requested but the peer exchange could not be set up')
        return crit

    def timed(self, step, steps, warmup, profile_every=0, sample_clocks=False):
        """W untimed + exactly K timed steps between barrier + synchronize, CUDA events on the
        launching stream, max over ranks.  Returns (ms_per_step, kernels, launches, clocks, last)."""
        torch, lib = self.torch, self.lib
        last = None
        for _ in range(max(warmup, 3)):
            last = step()
        self.barrier()
        sampler = None
        if sample_clocks and self.rank == 0:   # the line reports rank 0's GPU
            sampler = ClockSampler(self.local_rank)
            sampler.start()
        if profile_every:
            lib.profile_start()
        launches0 = lib.launch_count()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # no garbage-collector pauses inside the timed region: with one process per GPU a pause on
        # ANY rank stalls every rank at the next exchange
        gc.collect()
        gc_was = gc.isenabled()
        gc.disable()
        self.barrier()
        start.record()
        for i in range(steps):
            # per-kernel CUDA events on every `profile_every`-th step of the timed region
            # (bracketing every launch costs ~14 event records per step)
            if profile_every > 1:
                (lib.profile_resume if i % profile_every == 0 else lib.profile_pause)()
            last = step()
        stop.record()
        self.barrier()
        if gc_was:
            gc.enable()
        elapsed_ms = start.elapsed_time(stop)
        launches = lib.launch_count() - launches0
        kernels = lib.profile_stop() if profile_every else {}
        clocks = sampler.stop() if sampler else None
        return self.max_over_ranks(elapsed_ms) / steps, kernels, launches, clocks, last


def run_b200(args, out):
    import torch
    from b200det import synth, decode

    R = Runner(args)
    world, rank, dev = R.world, R.rank, R.dev
    B = local_batch(args, world)
    first = rank * B if args.scaling == 'strong' else 0
    seed = 100 if args.scaling == 'strong' else 100 + rank
    # the rank's shard of the global batch (per-image seeds: the same batch however it is sharded)
    preds = synth.make_retina_preds_sharded(first, B, SIZE, NUM_CLASSES, seed=seed, device=dev)
    ann = synth.make_annotations_sharded(first, B, MAX_GT, SIZE, NUM_CLASSES, seed=seed).to(dev)
    crit = R.make_criterion()
    dec = decode.RetinaDecoder(**synth.RETINA_KW)

    def step():
        with torch.no_grad():
            d = crit(preds, ann)
            r = dec(preds)
        return d, r

    # per-kernel CUDA events on every 4th step of the timed region: bracketing every launch of every
    # step costs ~0.02 ms per step (1.846 vs 1.829 ms at batch 256, tools/r02c_run9.sh)
    profile_every = args.profile_every or 4
    # The two calls share one sweep over cls when the decoder is called on the tensors the criterion
    # just read (b200det/_handoff.py; the library's default).  First the same loop with the hand-over
    # switched off -- every call sweeps for itself, r01 / early-r02 behaviour -- as `separate_sweeps`.
    from b200det import _handoff
    separate = None
    if _handoff.ENABLED and not args.no_separate:
        _handoff.ENABLED = False
        sms, sk, _, _, _ = R.timed(step, args.steps, args.warmup, profile_every=profile_every)
        _handoff.ENABLED = True
        _handoff.reset()
        separate = {
            'value': B * world / (sms / 1e3), 'unit': 'images/s', 'ms_per_step': sms,
            'kernels_ms': {k: round(v[1], 4) for k, v in sk.items()},
            'note': 'B200DET_HANDOFF=0: criterion and decoder each stream cls (2 x 4NC bytes per image)',
        }
    handoff0 = dict(_handoff.stats)
    ms_per_step, kernels, launches, clocks, (d, r) = R.timed(
        step, args.steps, args.warmup, profile_every=profile_every, sample_clocks=True)
    handoff = {k: _handoff.stats[k] - handoff0[k] for k in handoff0} if _handoff.ENABLED else None
    value = B * world / (ms_per_step / 1e3)
    status = crit.last_stats.get('exchange_status') if crit.last_stats else None
    exchange_status = int(status.item()) if status is not None else 0

    # ---- optional extension: one sweep over cls for loss + decode (b200det.fused.EvalStep) ----
    fused_info = None
    if not args.no_fused:
        from b200det import fused
        fstep = fused.EvalStep(crit, dec)
        fms = R.timed(lambda: fstep(preds, ann), args.steps, 3)[0]
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
        e2e = run_e2e(R, args, preds, ann, crit, dec)

    # ---- N > 1: sharded result == the reference's unsharded result, checked in this very run ----
    parity = None
    if R.distributed:
        parity = parity_check(R, crit)

    # ---- N > 1: round 1's weak-scaling definition beside the headline -----------------------
    weak = None
    if R.distributed and args.scaling == 'strong' and not args.no_weak:
        del preds
        torch.cuda.empty_cache()
        wB = GLOBAL_BATCH
        wpreds = synth.make_retina_preds(wB, SIZE, NUM_CLASSES, seed=100 + rank, device=dev)
        wann = synth.make_annotations(wB, MAX_GT, SIZE, NUM_CLASSES, seed=200 + rank).to(dev)

        def wstep():
            with torch.no_grad():
                return crit(wpreds, wann), dec(wpreds)

        wms, wk, _, _, _ = R.timed(wstep, max(10, args.steps // 4), 3, profile_every=1)
        weak = {
            'value': wB * world / (wms / 1e3), 'unit': 'images/s', 'ms_per_step': wms,
            'images_per_gpu': wB, 'global_batch': wB * world, 'scaling': 'weak',
            'kernels_ms': {k: round(v[1], 4) for k, v in wk.items()},
        }
        del wpreds
        torch.cuda.empty_cache()
        preds = None

    # ---- roofline of the dominant kernel ----------------------------------------------------
    peak, peak_src = hbm_peak()
    n, c = rows_per_image(), NUM_CLASSES
    alg = {'focal_loss': B * 4 * n * c, 'score_argmax': B * 4 * n * c}
    # The two sweeps move the same bytes and take ~44 % of the step each.  The focal sweep shares the
    # GPU with the assignment + sparse kernels on the helper stream (its CUDA-event window therefore
    # also covers their ALU work); the arg-max sweep has the GPU to itself: it is the one whose launch
    # duration is a clean kernel time, and the one reported as `roofline`.  The focal sweep's window
    # is in `roofline_focal_window`.
    # With the hand-over there is ONE sweep per step: the fused one (focal sum + arg-max keys), timed
    # under the arg-max sweep's id, with the assignment + sparse kernels beside it on the helper stream.
    dom = 'score_argmax' if 'score_argmax' in kernels else 'focal_loss'
    fused_sweep = 'focal_loss' not in kernels
    dom_ms = kernels[dom][1]
    achieved = alg[dom] / (dom_ms / 1e3) / 1e9
    loss_b, dec_b = algorithmic_bytes_per_image()
    step_gbs = B * (loss_b + dec_b) / (ms_per_step / 1e3) / 1e9
    traffic = ncu_traffic()
    # one launch of every kernel per step: what the step spends outside its kernels (launch gaps,
    # the decoder's host synchronisation, Python between the two calls)
    # (the assignment and the sparse losses run on the helper stream INSIDE the sweep's window when
    # the batch is large enough for the fork, losses.py: they are not on the critical path then)
    beside = ('assign', 'sparse_losses') if B * n * c >= (64 << 20) else ()
    kernel_sum = sum(v[1] for k, v in kernels.items() if k not in beside)

    line = {
        'metric': METRIC,
        'value': value,
        'unit': 'images/s',
        'n_gpus': world,
        'steps': args.steps,
        'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step,
        'higher_is_better': True,
        'scaling': args.scaling,
        'vs_baseline': None,
        'dtype': 'f32',
        'data': 'synthetic',
        'config': workload_config(B, world, args.scaling),
        'exchange': {'nccl': 'NCCL all-reduce of 4 doubles per step',
                     'p2p': '4 doubles per step over NVLink peer memory inside the reduction kernel',
                     'none': 'single GPU: no exchange'}[R.exchange],
        'exchange_status': exchange_status,
        'host_cpus_per_rank': R.cpus,
        'roofline': {
            'bound': 'hbm',
            'kernel': 'fused_sweep (fused_rows_tma_kernel: focal sum + arg-max keys from one read of cls; '
                      'assignment + sparse losses beside it on the helper stream)' if fused_sweep else dom,
            'achieved': achieved,
            'peak': peak,
            'unit': 'GB/s',
            'frac': achieved / peak,
            'traffic': (traffic.get('fused_sweep' if fused_sweep else dom) or 0) * B / 256 or None,
            'peak_source': peak_src,
            'algorithmic_bytes_per_launch': alg[dom],
            'kernel_ms': dom_ms,
        },
        'step_roofline': {
            'algorithmic_GBps': step_gbs,
            'frac_of_hbm_peak': step_gbs / peak,
            'note': 'whole step against the HBM peak for SURVEY 8d\'s algorithmic bytes (80.70 MB/image: '
                    'cls counted once per call, i.e. twice per step).  With the hand-over the step reads '
                    'cls once, so this fraction may exceed 1; see single_read_*',
            'single_read_bytes_per_image': loss_b + dec_b - 4 * n * c,
            'single_read_GBps': B * (loss_b + dec_b - 4 * n * c) / (ms_per_step / 1e3) / 1e9,
            'single_read_frac_of_hbm_peak': B * (loss_b + dec_b - 4 * n * c) / (ms_per_step / 1e3) / 1e9 / peak,
        },
        'handoff': handoff,
        'separate_sweeps': separate,
        'kernels_ms': {k: round(v[1], 4) for k, v in kernels.items()},
        'critical_path_kernels_ms': round(kernel_sum, 4),
        'outside_kernels_ms': round(ms_per_step - kernel_sum, 4),
        'clocks': clocks,
        'gpu_launches': launches,
        'e2e': e2e,
        'fused_eval_step': fused_info,
        'loss': {k: float(v.item()) for k, v in d.items()},
    }
    if parity is not None:
        line['parity_check'] = parity
    if weak is not None:
        line['weak_scaling'] = weak
    if rank == 0 and world == 1 and not args.no_configs:
        line['configs'] = run_configs(args)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ref = load_reference()
        steps = args.cpu_steps or 40   # ~11 s of CPU work for the reference classes
        v, dt, threads = time_cpu(ref, 'retina', 2, steps, 1)
        line['cpu_baseline'] = {
            'value': v,
            'unit': 'images/s',
            'cores': threads,
            'kind': ref[0],
            'sample': cpu_sample_note(ref[0], 2, steps, dt),
        }
    failed = exchange_status != 0 or (parity is not None and not parity['ok'])
    if rank == 0:
        out.emit(json.dumps(line))
    if R.distributed:
        R.dist.destroy_process_group()
    if failed:
        print('[bench] FAILED: sharded loss / labels differ from the unsharded batch, or the peer '
              'exchange timed out', file=sys.stderr)
        sys.exit(3)


def parity_check(R, crit):
    """Outside every timed region: a 4*N-image batch (per-image seeds: every rank can build all of
    it) is evaluated UNSHARDED by every rank (plain criterion: the reference's single-process
    semantics, losses.py:231-259) and SHARDED (this rank's 4 images, normaliser exchanged).  The
    sharded losses must equal the unsharded ones within 1e-5 relative, the labels bit for bit."""
    import torch
    from b200det import synth
    world, rank, dev, dist = R.world, R.rank, R.dev, R.dist
    per = 4
    full = synth.make_retina_preds_sharded(0, per * world, SIZE, NUM_CLASSES, seed=7, device=dev)
    fann = synth.make_annotations_sharded(0, per * world, MAX_GT, SIZE, NUM_CLASSES, seed=7).to(dev)
    mine = [[t[rank * per:(rank + 1) * per].contiguous() for t in grp] for grp in full]
    mann = fann[rank * per:(rank + 1) * per].contiguous()
    plain = R.make_criterion(sync=False)
    with torch.no_grad():
        want = plain(full, fann)
        got = crit(mine, mann)
    rel = 0.0
    for k in want:
        w, g = float(want[k].item()), float(got[k].item())
        rel = max(rel, abs(g - w) / max(abs(w), 1e-12))
    lab_full = plain.debug_assign(full, fann, exact=False)['labels'][rank * per:(rank + 1) * per]
    lab_mine = crit.debug_assign(mine, mann, exact=False)['labels']
    flags = torch.tensor([float(torch.equal(lab_full, lab_mine))], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    status = crit.last_stats.get('exchange_status') if crit.last_stats else None
    st = torch.tensor([float(status.item()) if status is not None else 0.0, rel],
                      dtype=torch.float64, device=dev)
    dist.all_reduce(st, op=dist.ReduceOp.MAX)
    rel = float(st[1].item())
    ok = rel <= 1e-5 and flags.item() > 0 and st[0].item() == 0
    return {
        'images': per * world,
        'sharded_vs_unsharded_rel': rel,
        'tolerance_rel': 1e-5,
        'labels_equal': bool(flags.item() > 0),
        'exchange_status': int(st[0].item()),
        'unsharded_loss': {k: float(v.item()) for k, v in want.items()},
        'sharded_loss': {k: float(v.item()) for k, v in got.items()},
        'ok': bool(ok),
    }


def _set_mempolicy(mode, node):
    """set_mempolicy(2) through libc's syscall(): page placement of what this thread touches next."""
    import ctypes
    libc = ctypes.CDLL(None, use_errno=True)
    if node is None:
        return libc.syscall(238, 0, None, 0)                     # MPOL_DEFAULT
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(238, mode, ctypes.byref(mask), 64)


def prefer_numa_node_of(bus_id):
    """Prefers the NUMA node the GPU hangs off for the pinned staging buffers allocated next (a
    cpuset that confines all ranks to one socket's CPUs would otherwise first-touch every rank's
    buffers on that socket, and half of the GPUs would read them across the socket link).  Returns
    the node, or None when it is unknown or the policy cannot be set."""
    try:
        bus = bus_id.decode() if isinstance(bus_id, bytes) else str(bus_id)
        bus = bus.lower()
        if len(bus.split(':')[0]) == 8:          # NVML prints an 8-digit PCI domain, sysfs has 4
            bus = bus[4:]
        with open(f'/sys/bus/pci/devices/{bus}/numa_node') as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        if _set_mempolicy(1, node) != 0:          # MPOL_PREFERRED
            return None
        return node
    except Exception:   # noqa: BLE001
        return None


def set_mempolicy_default():
    try:
        _set_mempolicy(0, None)
    except Exception:   # noqa: BLE001
        pass


def run_e2e(R, args, preds, ann, crit, dec):
    """Same step through the public classes, but starting from pinned host buffers every step."""
    torch = R.torch
    dev = R.dev
    B = args.e2e_batch or preds[0][0].shape[0]
    try:
        import psutil
        need = B * algorithmic_bytes_per_image()[0] * 1.1
        if psutil.virtual_memory().available < 3 * need * max(R.world, 1):
            B = max(8, B // 8)
    except Exception:
        pass
    # allocate the pinned staging buffers from the NUMA node next to this GPU (first touch), then
    # give the thread its old CPU mask back (the CPU baseline must keep all host cores)
    old_mask = os.sched_getaffinity(0)
    cpus = None
    numa = None
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(dev.index or 0)
        try:
            pynvml.nvmlDeviceSetCpuAffinity(handle)
            cpus = len(os.sched_getaffinity(0))
        except Exception as e:  # restricted cpuset: keep the CPU placement
            print(f'[bench] CPU affinity to the GPU skipped: {e}', file=sys.stderr)
        numa = prefer_numa_node_of(pynvml.nvmlDeviceGetPciInfo(handle).busId)
    except Exception as e:  # no NVML: keep the default placement
        print(f'[bench] NUMA binding skipped: {e}', file=sys.stderr)
    try:
        host = [[t[:B].cpu().pin_memory() for t in grp] for grp in preds]
        host_ann = ann[:B].cpu().pin_memory()
    finally:
        os.sched_setaffinity(0, old_mask)
        if numa is not None:
            set_mempolicy_default()
    h2d = sum(t.numel() * t.element_size() for grp in host for t in grp)
    h2d += host_ann.numel() * host_ann.element_size()
    d2h = 6 * B * 100 * 4 + 2 * 4

    def copies():
        p = [[t.to(dev, non_blocking=True) for t in grp] for grp in host]
        a = host_ann.to(dev, non_blocking=True)
        return p, a

    def step():
        with torch.no_grad():
            p, a = copies()
            d = crit(p, a)
            r = dec(p)
            vals = torch.stack([d['cls_loss'], d['reg_loss']]).cpu()   # D2H of the losses
        return vals, r

    def wall(fn, steps):
        fn()
        R.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        R.barrier()
        return R.max_over_ranks(time.perf_counter() - t0)

    dt = wall(step, args.e2e_steps)
    dt_copy = wall(copies, args.e2e_steps)
    return {
        'value': B * R.world * args.e2e_steps / dt,
        'unit': 'images/s',
        'h2d_bytes_per_step': h2d,
        'd2h_bytes_per_step': d2h,
        'images_per_gpu_per_step': B,
        'steps': args.e2e_steps,
        'h2d_only': {
            'GBps_per_gpu': h2d * args.e2e_steps / dt_copy / 1e9,
            'images_per_s': B * R.world * args.e2e_steps / dt_copy,
            'note': 'the same pinned-host -> HBM copies alone, all ranks at once: the ceiling of '
                    'this number (PCIe / host memory bandwidth), e2e / h2d_only = share of it reached',
        },
        'frac_of_h2d_only': (dt_copy / dt),
        'pinned_alloc_cpus': cpus,
        'pinned_alloc_numa_node': numa,
        'note': 'pinned host -> HBM copy of all head outputs + annotations, loss, decode+NMS, D2H '
                'of losses and detections, every step; PCIe-bound (129 kB of input per image-level '
                'row block: 40.3 MB per image against 80.7 MB of HBM traffic)',
    }


# --------------------------------------------------------------------------------------------
# BASELINE configs[0..3] + the training step of configs[4] (N = 1)
# --------------------------------------------------------------------------------------------
def time_config(name, kind, size, C, B, G, reps, box='GIoU', sigma=1.0, stages=('loss_fwd',
                'loss_fwd_bwd', 'decode_nms'), cpu=None):
    import torch
    from b200det import synth, losses, decode
    dev = torch.device('cuda')
    sizes = synth.pyramid_sizes(size)
    if kind == 'retina':
        preds = synth.make_retina_preds(B, size, C, seed=1, sigma=sigma, device=dev)
        crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type=box)
        dec = decode.RetinaDecoder(**synth.RETINA_KW)
        N, k = sum(p * p * 9 for p in sizes), 0
    else:
        preds = synth.make_fcos_preds(B, size, C, seed=1, sigma=sigma, device=dev)
        crit = losses.FCOSLoss(strides=synth.STRIDES, mi=synth.MI)
        dec = decode.FCOSDecoder(strides=synth.STRIDES)
        N, k = sum(p * p for p in sizes), 1
    ann = synth.make_annotations(B, G, size, C, seed=2).to(dev)
    loss_b, bwd_b, dec_b = stage_bytes(N, C, G, k)

    def fwd():
        with torch.no_grad():
            return crit(preds, ann)

    req = None

    def fwd_bwd():
        for grp in req:
            for t in grp:
                t.grad = None
        sum(crit(req, ann).values()).backward()

    def dec_fn():
        return dec(preds)

    def timed(fn, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    def eval_step():
        # the reference's evaluation step: both calls on the same head outputs (tools/scripts.py:733-740);
        # from the second iteration on the criterion's sweep hands the decoder its keys (_handoff.py)
        with torch.no_grad():
            d = crit(preds, ann)
        return d, dec(preds)

    pk = hbm_peak()[0]
    out = {'name': name, 'batch': B, 'rows_per_image': N, 'classes': C, 'max_gt': G}
    # stages are timed one after the other; the separate loss / decode stages come first, so no
    # decoder has asked for a hand-over yet when the criterion-only loop runs (it sweeps focal-only)
    for key, fn, nbytes in (('loss_fwd', fwd, loss_b), ('loss_fwd_bwd', fwd_bwd, loss_b + bwd_b),
                            ('decode_nms', dec_fn, dec_b),
                            ('eval_step', eval_step, loss_b + dec_b - 4 * N * C)):
        if key not in stages and not (key == 'eval_step' and 'loss_fwd' in stages and 'decode_nms' in stages):
            continue
        if key == 'loss_fwd_bwd':
            req = [[t.detach().requires_grad_(True) for t in grp] for grp in preds]
        ms = timed(fn)
        out[key] = {'ms': round(ms, 4), 'images_per_s': round(B / ms * 1e3, 1),
                    'frac_of_hbm_peak': round(B * nbytes / (ms * 1e-3) / 1e9 / pk, 4)}
        if key == 'eval_step':
            out[key]['note'] = ('criterion(preds, annots) + decoder(preds), cls streamed once (sweep hand-over); '
                                'fraction = bytes moved (loss_fwd + decode - 4NC) / time / HBM peak')
    from b200det import _handoff
    _handoff.reset()
    del preds, req
    torch.cuda.empty_cache()
    if cpu is not None:
        ref, kind_cfg, per_step, steps = cpu
        v, dt, threads = time_cpu(ref, kind_cfg, per_step, steps, 0, size=size, classes=C, max_gt=G)
        out['cpu_baseline'] = {'value': round(v, 3), 'unit': 'images/s', 'cores': threads,
                               'kind': ref[0], 'sample': f'{per_step} image(s) x {steps} step(s), '
                               f'{dt:.1f} s, stages: {kind_cfg}'}
    return out


def run_configs(args):
    """BASELINE.json configs[0..3], each stage timed with CUDA events (inputs resident in HBM), and
    the training step (loss forward + backward) of configs[4]; one bounded CPU sample of the
    reference per config."""
    reps = args.config_reps
    ref = None if args.no_cpu_baseline else load_reference()

    def cpu(kind_cfg, per_step=1, steps=2):
        return None if ref is None else (ref, kind_cfg, per_step, steps)

    res = [
        time_config('configs[0] RetinaDecoder + NMS, 800x800 C80 B=1', 'retina', 800, 80, 1, 100,
                    reps, stages=('decode_nms',), cpu=cpu('retina_decode', 1, 3)),
        time_config('configs[1] RetinaLoss(GIoU) 800x800 C80 B=16 G<=100', 'retina', 800, 80, 16, 100,
                    reps, stages=('loss_fwd', 'loss_fwd_bwd'), cpu=cpu('retina_loss', 2, 2)),
        time_config('configs[2] FCOSLoss + FCOSDecoder 800x800 C80 B=16', 'fcos', 800, 80, 16, 100,
                    reps, cpu=cpu('fcos', 2, 2)),
        time_config('configs[3] FCOS loss+decode Objects365 1024x1024 C365 B=32 G<=200', 'fcos',
                    1024, 365, 32, 200, reps, cpu=cpu('fcos', 1, 2)),
        time_config('configs[4] train_step: RetinaLoss(GIoU) forward + backward, B=256', 'retina',
                    800, 80, GLOBAL_BATCH, 100, max(5, reps // 4), stages=('loss_fwd_bwd',)),
    ]
    return res


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
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'],
                    help='strong: ONE global batch of 256 sharded over the GPUs (BASELINE configs[4], '
                         'SURVEY 8e); weak: 256 images per GPU (round 1)')
    ap.add_argument('--batch', type=int, default=0, help='images per GPU per step (overrides --scaling)')
    ap.add_argument('--e2e-batch', type=int, default=0)
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--cpu-steps', type=int, default=0)
    ap.add_argument('--config-reps', type=int, default=20)
    ap.add_argument('--ref-size', type=int, default=SIZE,
                    help='image size of the --impl reference sample (tests use a small one)')
    ap.add_argument('--sync', default='auto', choices=['auto', 'p2p', 'nccl'],
                    help='N > 1: how the loss normaliser crosses GPUs')
    ap.add_argument('--no-fused', action='store_true')
    ap.add_argument('--no-weak', action='store_true')
    ap.add_argument('--no-separate', action='store_true',
                    help='skip the extra timed loop with the sweep hand-over switched off')
    ap.add_argument('--no-pin', action='store_true', help='N > 1: do not give every rank its own CPUs')
    ap.add_argument('--no-configs', action='store_true')
    ap.add_argument('--profile-every', type=int, default=0)
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
