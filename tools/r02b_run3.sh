mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_handoff.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/b3_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/b3_tests.log
for b in 256 32; do
  echo "== tma batch $b overlap"; timeout 300 python tools/prof_eval_step.py --batch $b 2>&1 | tail -1
  echo "== tma batch $b serial"; B200DET_LOSS_OVERLAP=0 timeout 300 python tools/prof_eval_step.py --batch $b 2>&1 | tail -1
done
echo "== register-fed, serial"; B200DET_FUSED_NO_TMA=1 B200DET_LOSS_OVERLAP=0 timeout 300 python tools/prof_eval_step.py --batch 256 2>&1 | tail -1
