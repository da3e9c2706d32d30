python -m pytest tests -m gpu -x -q > gpurun_out/r02_t20.log 2>&1; tail -2 gpurun_out/r02_t20.log
for B in 256 32; do
B200DET_LOSS_OVERLAP=0 python bench.py --batch $B --steps 100 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_v20_noov_b$B.json 2> gpurun_out/r02_v20_noov_b$B.err
python bench.py --batch $B --steps 100 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_v20_b$B.json 2> gpurun_out/r02_v20_b$B.err
B200DET_ASSIGN_ANCHOR_CENTRIC=1 python bench.py --batch $B --steps 100 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_v20_old_b$B.json 2> gpurun_out/r02_v20_old_b$B.err
done
bash tools/ncu_assign.sh
