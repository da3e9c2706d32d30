# r02 (third session): raw-tile fused sweep (class counts that are not a multiple of 4): parity + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_handoff.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_gpu_vs_reference.py tests/test_config_dropin.py -m gpu -x -q > gpurun_out/c6_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/c6_tests.log
timeout 300 python tools/prof_cfg4_eval.py
timeout 300 python tools/prof_cfg4_eval.py --classes 80 --size 800 --batch 16 --gt 100
timeout 600 python bench.py --batch 32 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-fused --no-separate > gpurun_out/c6_cfg.json 2> gpurun_out/c6_cfg.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c6_cfg.json').read().strip().splitlines()[-1])
for c in d['configs']:
    print(c['name'], {k:(v['ms'], v['frac_of_hbm_peak']) for k,v in c.items() if isinstance(v,dict) and 'ms' in v})
PY
