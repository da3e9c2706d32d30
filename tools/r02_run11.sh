python -m pytest tests -m gpu -x -q > gpurun_out/r02_t11.log 2>&1; tail -2 gpurun_out/r02_t11.log
for i in 1 2; do
python bench.py --batch 32 --steps 300 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_v11_$i.json 2> gpurun_out/r02_v11_$i.err
done
python tools/host_overhead.py > gpurun_out/r02_host11.txt 2>&1; head -3 gpurun_out/r02_host11.txt
