#!/bin/bash
# Tuning aid (run on the GPU box): rebuilds the library with different compile-time knobs of the
# focal sweeps and prints the per-kernel times of a training step.  Leaves the default build.
set -u
for v in "1" "5" "6" "8"; do
  B200DET_NVCC_EXTRA="-DB200DET_FOCAL_GRAD_MINB=$v" python -c "import b200det; b200det._build.build(force=True)" > /dev/null 2>&1
  echo "minblocks=$v $(python tools/prof_train.py 2>&1 | tail -1)"
done
python -c "import b200det; b200det._build.build(force=True)" > /dev/null 2>&1
