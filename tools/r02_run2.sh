# r02 GPU run 2: bench at N=1 (headline + configs), the 32-image shard of the strong-scaling split, select A/B
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t2.log 2>&1; tail -3 gpurun_out/r02_t2.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_b1.json 2> gpurun_out/r02_b1.err; tail -c 600 gpurun_out/r02_b1.err; cat gpurun_out/r02_b1.json
python bench.py --batch 32 --steps 200 --warmup 10 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/r02_b32.json 2> gpurun_out/r02_b32.err; cat gpurun_out/r02_b32.json
python tools/prof_select.py > gpurun_out/r02_select_v2.json 2> gpurun_out/r02_select_v2.err
B200DET_SELECT_MINB=1 python tools/prof_select.py > gpurun_out/r02_select_v2_minb1.json 2> gpurun_out/r02_select_v2_minb1.err
python tools/host_overhead.py > gpurun_out/r02_host.txt 2>&1; cat gpurun_out/r02_host.txt
