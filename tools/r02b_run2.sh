# r02b run 2: TMA-fed fused sweep: parity tests that reach it + A/B against the register-fed kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_handoff.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/b2_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/b2_tests.log
for v in "B200DET_FUSED_NO_TMA=1" "B200DET_FUSED_MINB=4" "B200DET_FUSED_MINB=3"; do
  for b in 256 32; do
  echo "== $v batch $b"; env $v timeout 300 python tools/prof_eval_step.py --batch $b 2>&1 | tail -1
  done
done
