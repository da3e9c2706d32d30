#!/bin/bash
# Weak-scaling run on one 8-GPU box: N = 1, 2, 4, 8 back to back (gpurun --gpus 8 -- bash tools/scale_run.sh)
set -u
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 60 --warmup 5 --no-cpu-baseline --no-fused > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 60 --warmup 5 --no-fused > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
done
for n in 1 2 4 8; do
  python -c "
import json; d=json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1]); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), d['kernels_ms'], round(d['e2e']['value']), d['clocks']['sm_mhz'])"
done
