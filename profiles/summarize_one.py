#!/usr/bin/env python
"""Key ncu metrics of every kernel in one report -> a small text summary for profiles/.

    python profiles/summarize_one.py gpurun_out/logits_sweep.ncu-rep profiles/r01_logits_sweep.txt "note"
"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
note = sys.argv[3] if len(sys.argv) > 3 else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = [
    'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
    'dram__throughput.avg.pct_of_peak_sustained_elapsed',
    'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
    'launch__grid_size', 'launch__block_size', 'smsp__inst_executed.sum',
    'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
    'lts__t_sector_hit_rate.pct',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
]
ik = hdr.index('Kernel Name')
lines = [f'# {rep.split("/")[-1]}: ncu --set full --clock-control none (one launch; times under ncu are '
         'not bench values)']
if note:
    lines.append(f'# {note}')
for r in rows[2:]:
    lines.append(f'== {r[ik][:110]}')
    for name in WANT:
        if name in hdr:
            i = hdr.index(name)
            lines.append(f'   {name:<88s} {r[i]:>16s} {units[i]}')
open(out, 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:6]))
