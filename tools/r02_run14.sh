python -m pytest tests -m gpu -x -q > gpurun_out/r02_t14.log 2>&1; tail -3 gpurun_out/r02_t14.log
python tools/host_overhead.py > gpurun_out/r02_host14.txt 2>&1; head -3 gpurun_out/r02_host14.txt
B200DET_FASTPATH=0 python tools/host_overhead.py 2>&1 | head -3
for i in 1 2; do
python bench.py --batch 32 --steps 300 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_v14_$i.json 2> gpurun_out/r02_v14_$i.err
done
B200DET_FASTPATH=0 python bench.py --batch 32 --steps 300 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_v14_nofast.json 2> gpurun_out/r02_v14_nofast.err
python tools/prof_small.py > gpurun_out/r02_small14.json 2> gpurun_out/r02_small14.err
