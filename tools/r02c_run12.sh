# r02 (third session), 2 GPUs: driver-style N=2 line after the last changes (in-run parity check: sharded == unsharded)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c12_n2.json 2> gpurun_out/c12_n2.err; echo "N=2 rc=$?"; tail -c 300 gpurun_out/c12_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c12_n2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','kernels_ms','handoff','outside_kernels_ms','parity_check','exchange_status')})
print(d['separate_sweeps']['value'], d['weak_scaling']['value'], d['e2e']['value'])
PY
