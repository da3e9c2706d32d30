#!/bin/bash
# Tuning aid (GPU box): cache policy of the focal sweep's loads while the assignment runs beside it.
for v in 0 1; do
  B200DET_NVCC_EXTRA="-DB200DET_FOCAL_LOAD=$v" python -c "import b200det; b200det._build.build(force=True)" > /dev/null 2>&1
  for B in 256 32; do
    for ov in 1 0; do
    B200DET_LOSS_OVERLAP=$ov python bench.py --batch $B --steps 100 --warmup 5 --no-configs --no-cpu-baseline --no-e2e --no-fused 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('load=$v B=$B overlap=$ov ms', round(d['ms_per_step'],4), 'focal', d['kernels_ms']['focal_loss'], 'assign', d['kernels_ms']['assign'])"
    done
  done
done
python -c "import b200det; b200det._build.build(force=True)" > /dev/null 2>&1
