python -m pytest tests -m gpu -x -q > gpurun_out/r02_t24.log 2>&1; tail -2 gpurun_out/r02_t24.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_b24.json 2> gpurun_out/r02_b24.err
python tools/prof_train.py 2>&1 | tail -3
