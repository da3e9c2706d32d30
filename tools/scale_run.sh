#!/bin/bash
# The driver's scaling sequence on ONE 8-GPU box, back to back: N = 1, 2, 4, 8 with the driver's flags.
# gpurun --gpus 8 --timeout 1500 -- 'bash tools/scale_run.sh r02'
TAG=${1:-r02}
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${TAG}_scale_n1.json 2> gpurun_out/${TAG}_scale_n1.err; echo "N=1 rc=$?"
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$N \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err; echo "N=$N rc=$?"
done
