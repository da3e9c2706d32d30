# debug build (device-side capacity asserts) through the whole GPU suite, then the release build again
B200DET_NVCC_EXTRA=-DB200DET_DEBUG python -c "import b200det; b200det._build.build(force=True)" > /dev/null 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t16_debug.log 2>&1; tail -2 gpurun_out/r02_t16_debug.log
python -c "import b200det; b200det._build.build(force=True)" > /dev/null 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/r02_t16.log 2>&1; tail -2 gpurun_out/r02_t16.log
python tools/prof_select.py --reps 100 --only retina_b1,retina_b32 > gpurun_out/r02_select_v16.json 2> gpurun_out/r02_select_v16.err
