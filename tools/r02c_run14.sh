# r02 (third session): hand-over wish registered after a successful own decode: hand-over / fuzz / parity tests + a short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/i_tests.log
timeout 300 python bench.py --batch 256 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --no-fused --no-separate > gpurun_out/i_b256.json 2> gpurun_out/i_b256.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/i_b256.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],4), d['kernels_ms'], d['handoff'])
PY
