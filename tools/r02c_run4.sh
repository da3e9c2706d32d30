# r02 (third session): A/B of helper-stream kernels shaped to fit BESIDE three resident sweep CTAs
# (fused sweep: 3 x 256 threads x 80 registers = 61440 of 65536 registers per SM; leftover 4096 =
# 128 threads x 32 registers = 64 threads x 64 registers)
# variants/: assign.cu rebuilt with -DB200DET_TILE_THREADS=128 (t128), + -DB200DET_TILE_MINB=16 (t128r32),
# + -DB200DET_SPARSE_THREADS=64 (t128r32_s64), linked with the other objects of csrc/_obj (scratch, not kept)
mkdir -p gpurun_out
LIB=simpleaicv-pytorch-imagenet-coco-training_b200/libb200det.so
for v in default t128 t128r32 t128r32_s64; do
  cp variants/lib_$v.so $LIB
  for b in 256 32; do
    timeout 300 python bench.py --batch $b --steps 60 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --no-fused --no-separate > gpurun_out/c4_${v}_b$b.json 2> gpurun_out/c4_${v}_b$b.err; echo "$v b$b rc=$?"
  done
  B200DET_LOSS_OVERLAP=0 timeout 300 python tools/prof_eval_step.py --batch 256 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_vs_reference.py -m gpu -x -q > gpurun_out/c4_tests.log 2>&1; echo "tests(t128r32_s64) rc=$?"; tail -2 gpurun_out/c4_tests.log
cp variants/lib_default.so $LIB
python - <<'PY'
import json
for v in ('default','t128','t128r32','t128r32_s64'):
    for b in (256,32):
        d=json.loads(open(f'gpurun_out/c4_{v}_b{b}.json').read().strip().splitlines()[-1])
        print(v, b, round(d['ms_per_step'],4), d['kernels_ms'], d['clocks'] and d['clocks']['sm_mhz'])
PY
