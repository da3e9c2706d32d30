#!/usr/bin/env python
"""Turns the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

    python profiles/summarize.py r01

Reads gpurun_out/<tag>_full.ncu-rep (ncu --set full) and gpurun_out/<tag>_launches.csv (per-launch
gpu__time_duration) and writes profiles/<tag>_kernels.json, profiles/<tag>_launches.txt,
profiles/<tag>_hot_lines.txt and profiles/traffic.json (dram bytes per launch, read by bench.py)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
rep = os.path.join(ROOT, 'gpurun_out', f'{tag}_full.ncu-rep')

NAMES = {
    'focal_all_kernel': 'focal_loss',
    'focal_kernel': 'focal_loss_labelled',
    'retina_assign_kernel': 'retina_assign',
    'retina_assign_tile_kernel': 'retina_assign',
    'fused_rows': 'fused_sweep',
    'fcos_assign_kernel': 'fcos_assign',
    'sparse_loss_kernel': 'sparse_losses',
    'score_argmax_kernel': 'score_argmax',
    'select_nms_kernel': 'select_decode_nms',
    'loss_reduce_kernel': 'loss_reduce',
    'loss_reduce_exchange_kernel': 'loss_reduce',
    'loss_finish_kernel': 'loss_finish',
}


def short(name):
    for k, v in NAMES.items():
        if k in name:
            return v
    return name[:40]


raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = {
    'gpu__time_duration.sum': 'duration',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed': 'dram_pct_of_peak',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed': 'sm_pct_of_peak',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'occupancy_pct',
    'launch__registers_per_thread': 'registers',
    'launch__grid_size': 'grid',
    'launch__block_size': 'block',
    'smsp__inst_executed.sum': 'warp_instructions',
    'sm__inst_executed_pipe_fma.sum': 'fma_pipe_instructions',
    'l1tex__t_sector_hit_rate.pct': 'l1_hit_pct',
    'lts__t_sector_hit_rate.pct': 'l2_hit_pct',
}
scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9,
         'second': 1.0, 's': 1.0}
kernels = {}
traffic = {}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = short(d['Kernel Name'])
    rec = {'kernel': d['Kernel Name'][:90]}
    for m, k in want.items():
        if m in d:
            try:
                v = float(d[m].replace(',', ''))
            except ValueError:
                continue
            u = units[hdr.index(m)]
            rec[k] = v * scale[u] if u in scale else v
    if 'dram_read' in rec:
        rec['dram_bytes'] = rec['dram_read'] + rec.get('dram_write', 0.0)
        if rec.get('duration'):
            rec['dram_GBps_under_ncu'] = rec['dram_bytes'] / rec['duration'] / 1e9
        traffic[name] = rec['dram_bytes']
    kernels[name] = rec
json.dump(kernels, open(os.path.join(ROOT, 'profiles', f'{tag}_kernels.json'), 'w'), indent=1)
json.dump(traffic, open(os.path.join(ROOT, 'profiles', 'traffic.json'), 'w'), indent=1)

# launch list: mean device time per kernel and share of the step
lpath = os.path.join(ROOT, 'gpurun_out', f'{tag}_launches.csv')
if os.path.exists(lpath):
    text = [l for l in open(lpath) if l.startswith('"')]
    lr = list(csv.reader(io.StringIO(''.join(text))))
    h = lr[0]
    ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
    agg = {}
    for r in lr[1:]:
        if len(r) <= vi:
            continue
        t = float(r[vi].replace(',', '')) * scale.get(r[ui], 1e-6)
        agg.setdefault(short(r[ki]), []).append(t)
    ours = {k: v for k, v in agg.items() if k in NAMES.values()}
    mean = {k: sum(v) / len(v) for k, v in ours.items()}
    common = [k for k in ('retina_assign', 'fcos_assign', 'sparse_losses', 'loss_reduce', 'select_decode_nms') if k in mean]
    steps = {'hand-over step (cls read once: the default)': ['fused_sweep'] + common,
             'separate sweeps (B200DET_HANDOFF=0)': ['focal_loss', 'score_argmax'] + common}
    with open(os.path.join(ROOT, 'profiles', f'{tag}_launches.txt'), 'w') as f:
        f.write('# ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --batch 256 --steps 3 --warmup 3\n'
                '# (the separate-sweeps loop, then the hand-over loop; per-launch times are cold-cache and serialised:\n'
                '#  compare the SHARES with bench.py\'s kernels_ms, not the absolute times)\n')
        for title, names in steps.items():
            names = [k for k in names if k in mean]
            if not names or (names[0] not in ('fused_sweep', 'focal_loss')):
                continue
            step = sum(mean[k] for k in names)
            f.write(f'## {title}: {step * 1e6:.1f} us of kernels\n')
            f.write(f'{"kernel":24s} {"launches":>8s} {"mean_us":>10s} {"share_of_step":>14s}\n')
            for k in sorted(names, key=lambda k: -mean[k]):
                f.write(f'{k:24s} {len(ours[k]):8d} {mean[k] * 1e6:10.1f} {100 * mean[k] / step:13.1f}%\n')
        other = {k: v for k, v in agg.items() if k not in NAMES.values()}
        f.write('# other kernels in the same process (input generation with torch, memsets):\n')
        for k, v in sorted(other.items(), key=lambda kv: -sum(kv[1]))[:8]:
            f.write(f'#   {k:40s} {len(v):6d} launches, {sum(v) * 1e3:9.2f} ms total\n')

src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
tmp = os.path.join(ROOT, 'gpurun_out', f'{tag}_cs.csv')
open(tmp, 'w').write(src)
hot = subprocess.run([sys.executable, os.path.join(ROOT, 'profiles', 'ncu_lines.py'), tmp, '14'],
                     capture_output=True, text=True).stdout
open(os.path.join(ROOT, 'profiles', f'{tag}_hot_lines.txt'), 'w').write(hot)
for k, v in kernels.items():
    print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items() if a != 'kernel'})
