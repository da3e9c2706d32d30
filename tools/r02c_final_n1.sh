# r02 final (third session): full GPU suite, smoke(), driver-style bench line + reference arm, launch list and
# full ncu capture of the final code
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/f_smoke.log
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_n1.json 2> gpurun_out/f_n1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/f_n1.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; echo "ref rc=$?"
SKIP=30 COUNT=17 SKIP_SELECT=1 timeout 900 bash profiles/capture.sh r02f; echo "capture rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/f_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernels_ms','handoff','critical_path_kernels_ms','outside_kernels_ms','clocks','gpu_launches')})
print('  separate', d['separate_sweeps']['value'], d['separate_sweeps']['ms_per_step'])
print('  roofline', d['roofline'])
print('  e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'])
for c in d['configs']:
    print(c['name'], {k:(v['ms'], v['frac_of_hbm_peak']) for k,v in c.items() if isinstance(v,dict) and 'ms' in v})
PY
