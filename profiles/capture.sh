#!/bin/bash
# Captures the profiles summarised in profiles/ (run on a B200 through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh r01'
# 1) plain run (must exit 0), 2) launch list with per-launch device time, 3) ncu --set full of
# the five hot kernels at the benchmark batch.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --batch 256 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
$CMD > /dev/null 2>&1 || exit 1
# skip the 3 warm-up + first timed step (6 matching kernels per step), then take one step
ncu --set full --clock-control none --import-source on \
    -k regex:"focal_all_kernel|retina_assign_kernel|sparse_loss_kernel|score_argmax_kernel|select_nms_kernel|loss_reduce_kernel" \
    -s 24 -c 6 -o $OUT/${TAG}_full $CMD > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > $OUT/${TAG}_smi.csv
