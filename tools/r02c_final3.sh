# r02 final check (third session) after the tile-size switch moved to 4 images: full GPU suite + smoke + small-batch criterion times
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/h_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/h_smoke.log
timeout 200 python tools/prof_loss_small.py
