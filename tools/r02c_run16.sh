# r02 (third session): RetinaFace pair through the hand-over (new GPU test) + the hand-over tests
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_handoff.py -m gpu -x -q 2>&1 | tail -3
