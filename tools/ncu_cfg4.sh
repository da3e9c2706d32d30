# ncu capture of the two class-count-not-a-multiple-of-4 sweeps at BASELINE configs[3] (C = 365)
out=${1:-gpurun_out/cfg4}
python tools/prof_cfg4.py > $out.plain.log 2>&1 || { tail -5 $out.plain.log; exit 1; }
tail -1 $out.plain.log
ncu --set full --clock-control none --import-source on -k regex:"score_argmax_raw_kernel|focal_kernel|focal_all_kernel" \
    -s 9 -c 3 -o $out -f python tools/prof_cfg4.py > $out.log 2>&1
tail -2 $out.log
