# r02 GPU run 6 (2 GPUs): strong-scaling bench line with parity check, p2p check under torchrun
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r02_n2.json 2> gpurun_out/r02_n2.err; echo rc=$?; tail -c 1500 gpurun_out/r02_n2.err; cat gpurun_out/r02_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/p2p_check.py > gpurun_out/r02_p2p.json 2> gpurun_out/r02_p2p.err; echo rc=$?; tail -c 800 gpurun_out/r02_p2p.err; cat gpurun_out/r02_p2p.json
python -m pytest tests -m gpu -x -q -k "peer_exchange" 2>&1 | tail -3
