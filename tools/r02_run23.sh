python -m pytest tests/test_gpu_fuzz.py -m gpu -x -q -k tile_assignment > gpurun_out/r02_t23.log 2>&1; tail -12 gpurun_out/r02_t23.log
