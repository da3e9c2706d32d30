python -m pytest tests -m gpu -x -q > gpurun_out/r02_t17.log 2>&1; tail -2 gpurun_out/r02_t17.log
python bench.py --steps 30 --warmup 5 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r02_b17.json 2> gpurun_out/r02_b17.err
python bench.py --batch 32 --steps 300 --warmup 10 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r02_b17_32.json 2> gpurun_out/r02_b17_32.err
