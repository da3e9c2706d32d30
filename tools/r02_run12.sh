for S in 1 2 4 8; do
B200DET_SELECT_SLICES=$S python tools/prof_select.py --reps 100 --only retina_b16,retina_b64,retina_b128,fcos_1024_c365_b32 > gpurun_out/r02_sl_$S.json 2> gpurun_out/r02_sl_$S.err
done
