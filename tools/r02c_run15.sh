# r02 (third session): the fused sweep as K PDL-chained launches with the helper stream's kernels issued after the
# first chunk (B200DET_SWEEP_CHUNKS), helper stream at high (1) / normal (0) priority (B200DET_SIDE_PRIORITY)
mkdir -p gpurun_out
run() {  # $1 = tag, $2 = batch, rest = env
  tag=$1; b=$2; shift 2
  env "$@" timeout 200 python bench.py --batch $b --steps 40 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --no-fused --no-separate --profile-every 1000 > gpurun_out/c15_${tag}_b$b.json 2> gpurun_out/c15_${tag}_b$b.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/c15_${tag}_b$b.json').read().strip().splitlines()[-1])
    print('$tag', $b, round(d['ms_per_step'],4), d['kernels_ms'], d['loss'])
except Exception as e:
    print('$tag', $b, 'FAILED', e, open('gpurun_out/c15_${tag}_b$b.err').read()[-300:])
PY
}
run default 256 A=1
run k8p0 256 B200DET_SWEEP_CHUNKS=8 B200DET_SIDE_PRIORITY=0
run k8p1 256 B200DET_SWEEP_CHUNKS=8 B200DET_SIDE_PRIORITY=1
run k4p0 256 B200DET_SWEEP_CHUNKS=4 B200DET_SIDE_PRIORITY=0
run k16p0 256 B200DET_SWEEP_CHUNKS=16 B200DET_SIDE_PRIORITY=0
run default 32 A=1
run k8p0 32 B200DET_SWEEP_CHUNKS=8 B200DET_SIDE_PRIORITY=0
run k4p0 32 B200DET_SWEEP_CHUNKS=4 B200DET_SIDE_PRIORITY=0
B200DET_SWEEP_CHUNKS=8 B200DET_SIDE_PRIORITY=0 timeout 300 python -m pytest tests/test_gpu_handoff.py -m gpu -x -q 2>&1 | tail -1
