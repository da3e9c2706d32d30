# r02 (third session): GPU suite, driver-style bench line, launch list + full ncu capture of the final code
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c1_tests.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c1_n1.json 2> gpurun_out/c1_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/c1_n1.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/c1_ref.json 2> gpurun_out/c1_ref.err; echo "ref rc=$?"
SKIP=30 COUNT=17 SKIP_SELECT=1 timeout 900 bash profiles/capture.sh r02c; echo "capture rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c1_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernels_ms','handoff','outside_kernels_ms','clocks')})
print('  separate', d['separate_sweeps'])
print('  roofline', d['roofline'])
print('  e2e', d['e2e'])
print('  fused', d['fused_eval_step'])
PY
