python -m pytest tests -m gpu -x -q > gpurun_out/r02_t22.log 2>&1; tail -2 gpurun_out/r02_t22.log
for B in 256 32; do
python bench.py --batch $B --steps 100 --warmup 10 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r02_v22_b$B.json 2> gpurun_out/r02_v22_b$B.err
done
