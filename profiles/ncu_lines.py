#!/usr/bin/env python
"""Summarises `ncu --page source --csv --print-source cuda,sass` output: per kernel, the CUDA source
lines ranked by executed warp instructions, with stall samples.  Usage: ncu_lines.py cs.csv [top]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
kernel = None
fname = '?'
hdr = None
data = {}
for r in rows:
    if not r:
        continue
    if r[0] == 'File Name':
        fname = r[1].split('/')[-1]
        continue
    if r[0] in ('Function Name', 'Kernel Name'):
        kernel = r[1][:70]
        data.setdefault(kernel, [])
        hdr = None
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or kernel is None:
        continue
    if r[0] != '':  # a CUDA source line summary row
        d = dict(zip(hdr, r))
        try:
            inst = int(d['Instructions Executed'])
            samp = int(d['# Samples'])
        except Exception:
            continue
        data[kernel].append((inst, samp, fname[:10] + ':' + r[0], r[1].strip()[:100]))
for k, v in data.items():
    tot_i = sum(x[0] for x in v) or 1
    tot_s = sum(x[1] for x in v) or 1
    print(f'== {k}: {tot_i} warp-instr, {tot_s} samples')
    for inst, samp, line, src in sorted(v, reverse=True)[:top]:
        print(f'  {100*inst/tot_i:5.1f}% inst {100*samp/tot_s:5.1f}% samp  {line:>15}: {src}')
