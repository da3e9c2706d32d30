# ncu capture of the training-path sweep (label-aware focal kernel that writes d(cls))
out=${1:-gpurun_out/train_sweep}
python tools/prof_train.py > $out.plain.log 2>&1 || { tail -5 $out.plain.log; exit 1; }
tail -1 $out.plain.log
ncu --set full --clock-control none --import-source on -k regex:focal_kernel -s 4 -c 1 -o $out -f \
    python tools/prof_train.py > $out.log 2>&1
tail -2 $out.log
