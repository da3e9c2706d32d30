python -m pytest tests -m gpu -x -q > gpurun_out/r02_t13.log 2>&1; tail -2 gpurun_out/r02_t13.log
python tools/prof_select.py --reps 100 --only retina_b64,retina_b128,retina_b256 > gpurun_out/r02_pf1.json 2> gpurun_out/r02_pf1.err
B200DET_SELECT_NO_PREFILTER=1 python tools/prof_select.py --reps 100 --only retina_b64,retina_b128,retina_b256 > gpurun_out/r02_pf0.json 2> gpurun_out/r02_pf0.err
python bench.py --steps 30 --warmup 5 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r02_b13.json 2> gpurun_out/r02_b13.err
python bench.py --batch 32 --steps 300 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_b13_32.json 2> gpurun_out/r02_b13_32.err
