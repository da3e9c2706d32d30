python -m pytest tests -m gpu -x -q > gpurun_out/r02_t18.log 2>&1; tail -5 gpurun_out/r02_t18.log
for B in 256 32; do
python bench.py --batch $B --steps 100 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_v18_b$B.json 2> gpurun_out/r02_v18_b$B.err
B200DET_ASSIGN_ANCHOR_CENTRIC=1 python bench.py --batch $B --steps 100 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_v18_old_b$B.json 2> gpurun_out/r02_v18_old_b$B.err
B200DET_LOSS_OVERLAP=0 python bench.py --batch $B --steps 100 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_v18_noov_b$B.json 2> gpurun_out/r02_v18_noov_b$B.err
done
