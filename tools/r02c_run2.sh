# r02 (third session), 2 GPUs: peer exchange check + driver-style N=2 line with the sweep hand-over
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/p2p_check.py > gpurun_out/c2_p2p.json 2> gpurun_out/c2_p2p.err; echo "p2p rc=$?"; tail -c 1500 gpurun_out/c2_p2p.json
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpu or peer or p2p or exchange" > gpurun_out/c2_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/c2_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c2_n2.json 2> gpurun_out/c2_n2.err; echo "N=2 rc=$?"; tail -c 400 gpurun_out/c2_n2.err
timeout 300 python bench.py --batch 32 --steps 100 --warmup 5 --no-e2e --no-cpu-baseline --no-configs > gpurun_out/c2_b32.json 2> gpurun_out/c2_b32.err; echo "b32 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/c2_n2.json','gpurun_out/c2_b32.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, {k:d.get(k) for k in ('value','ms_per_step','kernels_ms','handoff','outside_kernels_ms','parity_check','exchange','loss')})
    print('  separate', d['separate_sweeps'])
    print('  weak', d.get('weak_scaling'))
    print('  e2e', d['e2e'] and d['e2e']['value'], 'fused', d['fused_eval_step'])
PY
