# r02 GPU run 1: the cluster select kernel -- parity suite + phase breakdown, A/B vs single CTA
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t1.log 2>&1; tail -15 gpurun_out/r02_t1.log
timeout 300 python tools/prof_select.py > gpurun_out/r02_select_cluster.json 2> gpurun_out/r02_select_cluster.err; cat gpurun_out/r02_select_cluster.json
B200DET_SELECT_SLICES=1 timeout 300 python tools/prof_select.py > gpurun_out/r02_select_s1.json 2> gpurun_out/r02_select_s1.err; cat gpurun_out/r02_select_s1.json
B200DET_SELECT_BITONIC=1 timeout 300 python tools/prof_select.py --reps 100 > gpurun_out/r02_select_bitonic.json 2> gpurun_out/r02_select_bitonic.err; cat gpurun_out/r02_select_bitonic.json
