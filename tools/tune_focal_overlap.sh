#!/bin/bash
# Tuning aid (run on the GPU box): the label-free focal sweep's loads in flight / batches per CTA with
# the assignment + sparse kernels running beside it.  Leaves the default build.
set -u
for v in "4 2" "8 1" "8 2" "6 2" "4 4" "2 4"; do
  set -- $v
  B200DET_NVCC_EXTRA="-DB200DET_FOCAL_UNROLL=$1 -DB200DET_FOCAL_BATCHES=$2" python -c "import b200det; b200det._build.build(force=True)" > /dev/null 2>&1
  for B in 256 32; do
    python bench.py --batch $B --steps 100 --warmup 5 --no-configs --no-cpu-baseline --no-e2e --no-fused 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('unroll=$1 batches=$2 B=$B ms', round(d['ms_per_step'],4), 'focal', d['kernels_ms']['focal_loss'], 'argmax', d['kernels_ms']['score_argmax'])"
  done
done
python -c "import b200det; b200det._build.build(force=True)" > /dev/null 2>&1
