# r02b run 1: GPU tests of the new fused sweep + hand-over, fused-kernel A/B, first bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/b1_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/b1_tests.log
for v in "B200DET_FUSED_OLD=1" "B200DET_FUSED_MINB=4" "B200DET_FUSED_MINB=5" "B200DET_FUSED_MINB=6"; do
  echo "== $v"; env $v timeout 300 python tools/prof_eval_step.py --batch 256 2>&1 | tail -1
done
for v in "B200DET_FUSED_MINB=4" "B200DET_FUSED_MINB=5"; do
  echo "== $v batch 32"; env $v timeout 300 python tools/prof_eval_step.py --batch 32 2>&1 | tail -1
done
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/b1_n1.json 2> gpurun_out/b1_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/b1_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b1_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','kernels_ms','handoff')})
print('separate', d['separate_sweeps'])
print('roofline', d['roofline'])
print('fused', d['fused_eval_step'])
for c in d.get('configs',[]): print(c['name'], {k:v for k,v in c.items() if isinstance(v,dict) and 'ms' in v})
PY
