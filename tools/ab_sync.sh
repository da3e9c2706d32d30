# A/B at N GPUs: normaliser exchange over peer memory (p2p) vs NCCL.  usage: bash tools/ab_sync.sh N
N=${1:-2}
for s in p2p nccl p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 60 --warmup 5 --no-e2e --no-cpu-baseline --sync $s 2> gpurun_out/ab_sync_$s.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('sync=$s', 'n_gpus', d['n_gpus'], round(d['value'],1), 'img/s', round(d['ms_per_step'],4), 'ms;', d['config']['parallelism'], d['kernels_ms'])"
done
