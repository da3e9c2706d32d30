mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b4_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/b4_tests.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/b4_n1.json 2> gpurun_out/b4_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/b4_n1.err
timeout 300 python bench.py --batch 32 --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/b4_b32.json 2> gpurun_out/b4_b32.err; echo "bench32 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/b4_n1.json','gpurun_out/b4_b32.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k:d[k] for k in ('value','ms_per_step','kernels_ms','handoff','outside_kernels_ms')})
    print('  separate', d['separate_sweeps'])
    print('  roofline', {k:d['roofline'][k] for k in ('achieved','frac','kernel_ms')}, 'fused', d['fused_eval_step'] and d['fused_eval_step']['ms_per_step'])
PY
