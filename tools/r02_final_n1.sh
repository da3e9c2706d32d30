# the driver's own N = 1 invocation + the reference arm
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_ref.json 2> gpurun_out/r02_final_ref.err; echo rc=$?
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_final_n1.json 2> gpurun_out/r02_final_n1.err; echo rc=$?; tail -c 500 gpurun_out/r02_final_n1.err
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
