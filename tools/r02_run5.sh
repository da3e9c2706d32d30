# r02 GPU run 5: PDL select launch + two-phase loss call
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t5.log 2>&1; tail -3 gpurun_out/r02_t5.log
run() { tag=$1; shift; env "$@" python bench.py --batch 32 --steps 200 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_v5_$tag.json 2> gpurun_out/r02_v5_$tag.err; }
run base X=1
run nopdl B200DET_NO_PDL=1
run noprof X=1 
python tools/prof_select.py --only retina_b1,retina_b32,fcos_b16 > gpurun_out/r02_select_v5.json 2> gpurun_out/r02_select_v5.err
python tools/host_overhead.py > gpurun_out/r02_host5.txt 2>&1; head -3 gpurun_out/r02_host5.txt
