# r02 (third session): new full-size hand-over tests; cost of bracketing every launch with events in the timed region
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_handoff.py -m gpu -x -q > gpurun_out/c9_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/c9_tests.log
for pe in 1 4 1 4; do
timeout 300 python bench.py --batch 256 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --no-fused --no-separate --profile-every $pe > gpurun_out/c9_pe$pe.json 2> gpurun_out/c9_pe$pe.err; echo "pe=$pe rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/c9_pe$pe.json').read().strip().splitlines()[-1])
print($pe, round(d['ms_per_step'],4), d['kernels_ms'], d['outside_kernels_ms'])
PY
done
