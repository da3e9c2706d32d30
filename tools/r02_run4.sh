# r02 GPU run 4: how the side-stream assignment shares the SMs with the focal sweep (batch 32)
run() { tag=$1; shift; env "$@" python bench.py --batch 32 --steps 200 --warmup 10 --no-configs --no-cpu-baseline --no-e2e --no-fused > gpurun_out/r02_v4_$tag.json 2> gpurun_out/r02_v4_$tag.err; }
run base X=1
run prio0 B200DET_SIDE_PRIORITY=0
run chunk8 B200DET_ASSIGN_CHUNK=8
run chunk4 B200DET_ASSIGN_CHUNK=4
run chunk2 B200DET_ASSIGN_CHUNK=2
run chunk4p0 B200DET_ASSIGN_CHUNK=4 B200DET_SIDE_PRIORITY=0
