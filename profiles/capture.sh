#!/bin/bash
# Captures the profiles summarised in profiles/ (run on a B200 through gpurun from the repo root):
#   gpurun --timeout 1500 -- 'bash profiles/capture.sh r02'
# 1) plain run (must exit 0), 2) launch list with per-launch device time, 3) ncu --set full of
# the five hot kernels at the benchmark batch.  Numbers printed under ncu are never bench values.
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --batch 256 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-fused"
$CMD > $OUT/${TAG}_plain.json 2> $OUT/${TAG}_plain.err || { echo "plain run failed"; tail -5 $OUT/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
$CMD > /dev/null 2>&1 || exit 1
# bench.py first times the loop with the hand-over off (6 kernels per step: 6 x 6 = 36 launches),
# then with it on (5 per step).  Skip into the last separate-sweeps step and take it plus one hand-over step.
ncu --set full --clock-control none --import-source on \
    -k regex:"fused_rows|focal_all_kernel|retina_assign|sparse_loss_kernel|score_argmax_kernel|select_nms_kernel|loss_reduce" \
    -s ${SKIP:-30} -c ${COUNT:-11} -o $OUT/${TAG}_full $CMD > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,clocks_event_reasons.active --format=csv > $OUT/${TAG}_smi.csv

[ -n "${SKIP_SELECT:-}" ] && exit 0
# 4) the select kernel at batch 1 (BASELINE configs[0], latency-bound): launch list of one decoder
#    call and a full capture of the cluster kernel
cat > /tmp/dec1.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from b200det import synth, decode
B = int(sys.argv[1])
preds = synth.make_retina_preds(B, 800, 80, seed=1, device='cuda')
dec = decode.RetinaDecoder(**synth.RETINA_KW)
for _ in range(6): dec(preds)
torch.cuda.synchronize()
PY
for B in 1 32; do
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"select|score_argmax" \
    -s 8 -c 4 --csv --log-file $OUT/${TAG}_select_b${B}_launches.csv python /tmp/dec1.py $B > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"select_nms_kernel" -s 4 -c 1 \
    -o $OUT/${TAG}_select_b1 python /tmp/dec1.py 1 > $OUT/${TAG}_ncu_select.log 2>&1
tail -2 $OUT/${TAG}_ncu_select.log
