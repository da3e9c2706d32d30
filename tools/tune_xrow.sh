# A/B of the XROW training sweep's residency (CTAs per SM) at BASELINE configs[3]; rebuilds focal.o only
P=simpleaicv-pytorch-imagenet-coco-training_b200
for m in 6 5; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
       -DB200DET_FOCAL_XROW_MINB=$m -c $P/csrc/focal.cu -o $P/csrc/_obj/focal.o || exit 1
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $P/libb200det.so $P/csrc/_obj/*.o || exit 1
  echo "MINB=$m $(python tools/prof_cfg4.py 2>&1 | tail -1)"
done
