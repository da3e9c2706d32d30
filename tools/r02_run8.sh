# r02 GPU run 8 (2 GPUs): graph-replay check of the peer exchange + new GPU tests
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/p2p_check.py > gpurun_out/r02_p2p_graph.json 2> gpurun_out/r02_p2p_graph.err; echo rc=$?; tail -c 1500 gpurun_out/r02_p2p_graph.err; cat gpurun_out/r02_p2p_graph.json
python -m pytest tests -m gpu -x -q -k "nan or second_backward or peer_exchange" 2>&1 | tail -5
