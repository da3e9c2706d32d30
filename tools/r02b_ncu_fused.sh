# ncu of the fused sweep at batch 256: full set + source.  $1 = tag, rest = env assignments
mkdir -p gpurun_out
TAG=$1; shift
for v in "$@"; do export $v; done
export B200DET_LOSS_OVERLAP=0
timeout 300 python tools/prof_eval_step.py --batch 256 --iters 10 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fused_rows" -s 4 -c 1 \
    -f -o gpurun_out/r02b_fused_$TAG python tools/prof_eval_step.py --batch 256 --iters 3 > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log
