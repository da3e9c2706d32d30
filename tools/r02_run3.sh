# r02 GPU run 3: loss forward with assignment + sparse losses on a side stream beside the focal sweep
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t3.log 2>&1; tail -3 gpurun_out/r02_t3.log
for B in 32 256; do
  python bench.py --batch $B --steps 100 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_ov1_b$B.json 2> gpurun_out/r02_ov1_b$B.err
  B200DET_LOSS_OVERLAP=0 python bench.py --batch $B --steps 100 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_ov0_b$B.json 2> gpurun_out/r02_ov0_b$B.err
done
B200DET_SELECT_SLICES=8 python bench.py --batch 32 --steps 100 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_ov1_b32_s8.json 2> gpurun_out/r02_ov1_b32_s8.err
python bench.py --batch 32 --steps 100 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused --profile-every 1000 > gpurun_out/r02_ov1_b32_noprof.json 2> gpurun_out/r02_ov1_b32_noprof.err
python tools/prof_small.py > gpurun_out/r02_small.json 2> gpurun_out/r02_small.err
