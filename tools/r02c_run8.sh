# r02 (third session): trimmed ring kernel: parity of the hand-over tests + timing with / without overlap
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_handoff.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/c8_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/c8_tests.log
timeout 300 python tools/prof_cfg4_eval.py
B200DET_LOSS_OVERLAP=0 timeout 300 python tools/prof_cfg4_eval.py
