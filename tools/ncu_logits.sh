# ncu capture of the logits sweep (tools/bench_logits.py): full set, one launch after warm-up
out=${1:-gpurun_out/logits_sweep}
ncu --set full --clock-control none --import-source on -k regex:logits_sweep -s 3 -c 1 -o $out -f \
    python tools/bench_logits.py --batch 128 --iters 2 > $out.log 2>&1
tail -2 $out.log
