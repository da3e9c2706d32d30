# r02 (third session): the ring-fed raw fused sweep alone (no helper-stream overlap) + ncu full capture
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fused_raw_ring" -s 3 -c 1 -f -o gpurun_out/r02c_raw_ring env B200DET_LOSS_OVERLAP=0 python tools/prof_cfg4_eval.py --iters 3 > gpurun_out/c7_ncu.log 2>&1
tail -2 gpurun_out/c7_ncu.log
