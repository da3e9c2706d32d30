# r02 (third session): one tile-assignment CTA per SM (shared-memory pad) beside the sweep; side-24 tiles forced (parity)
mkdir -p gpurun_out
run() {  # $1 = tag, $2 = batch, rest = env
  tag=$1; b=$2; shift 2
  env "$@" timeout 300 python bench.py --batch $b --steps 60 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --no-fused --no-separate > gpurun_out/c11_${tag}_b$b.json 2> gpurun_out/c11_${tag}_b$b.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/c11_${tag}_b$b.json').read().strip().splitlines()[-1])
print('$tag', $b, round(d['ms_per_step'],4), d['kernels_ms'])
PY
}
for b in 256 32 16; do
  run default $b A=1
  run pad106 $b B200DET_TILE_MIN_SMEM_KB=106
  run pad140 $b B200DET_TILE_MIN_SMEM_KB=140
done
B200DET_TILE_SMALL_BATCH=100000 B200DET_TILE_SIDE_SMALL=24 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_vs_reference.py tests/test_gpu_handoff.py -m gpu -x -q > gpurun_out/c11_tests.log 2>&1; echo "tests(side24 everywhere) rc=$?"; tail -2 gpurun_out/c11_tests.log
