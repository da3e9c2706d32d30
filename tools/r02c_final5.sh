# r02 last check (third session) of the library as it ships: full GPU suite + smoke()
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/k_tests.log 2>&1; echo "tests rc=$?"; tail -1 gpurun_out/k_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
