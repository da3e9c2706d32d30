# r02 final record (third session): driver-style N=1 line + reference arm of the final library
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/j_n1.json 2> gpurun_out/j_n1.err; echo "bench rc=$?"; tail -c 200 gpurun_out/j_n1.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/j_ref.json 2> gpurun_out/j_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/j_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernels_ms','handoff','critical_path_kernels_ms','outside_kernels_ms','clocks','gpu_launches')})
print('  separate', d['separate_sweeps']['value'], d['separate_sweeps']['ms_per_step'], 'fused', d['fused_eval_step']['ms_per_step'])
print('  roofline', d['roofline']['frac'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'])
for c in d['configs']:
    print(c['name'], {k:(v['ms'], v['frac_of_hbm_peak']) for k,v in c.items() if isinstance(v,dict) and 'ms' in v})
PY
