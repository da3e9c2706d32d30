# r02 GPU run 9 (8 GPUs): strong scaling at N=8 and N=4
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/r02_n$N.json 2> gpurun_out/r02_n$N.err; echo rc=$?; tail -c 600 gpurun_out/r02_n$N.err
done
