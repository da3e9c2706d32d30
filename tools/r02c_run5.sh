# r02 (third session), 8 GPUs: strong-scaling lines at N = 8 and N = 4 with the sweep hand-over (driver flags)
mkdir -p gpurun_out
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/c5_n$N.json 2> gpurun_out/c5_n$N.err; echo "N=$N rc=$?"; tail -c 300 gpurun_out/c5_n$N.err
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 8 --steps 100 --warmup 5 --no-e2e --no-weak --no-fused > gpurun_out/c5_n8_100.json 2> gpurun_out/c5_n8_100.err; echo "N=8 (100 steps) rc=$?"
python - <<'PY'
import json
for f in ('c5_n8','c5_n4','c5_n8_100'):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, {k:d.get(k) for k in ('value','ms_per_step','kernels_ms','handoff','outside_kernels_ms','exchange')})
    print('  parity', d.get('parity_check') and {k:d['parity_check'][k] for k in ('sharded_vs_unsharded_rel','labels_equal','exchange_status','ok')})
    print('  separate', d['separate_sweeps'] and (d['separate_sweeps']['value'], d['separate_sweeps']['ms_per_step']))
    print('  weak', d.get('weak_scaling') and (d['weak_scaling']['value'], d['weak_scaling']['ms_per_step']))
    print('  e2e', d['e2e'] and d['e2e']['value'], 'fused', d['fused_eval_step'] and d['fused_eval_step']['ms_per_step'])
PY
