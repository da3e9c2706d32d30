set -u
for v in "4 2" "8 1" "4 4" "2 4" "8 2"; do
  set -- $v
  B200DET_NVCC_EXTRA="-DB200DET_FOCAL_UNROLL=$1 -DB200DET_FOCAL_BATCHES=$2" python -c "import b200det; b200det._build.build(force=True)" > /dev/null 2>&1
  python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 $2', round(d['value']), d['kernels_ms']['focal_loss'], d['kernels_ms']['score_argmax'])"
done
python -c "import b200det; b200det._build.build(force=True)" > /dev/null 2>&1
