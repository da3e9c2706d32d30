# r02 GPU run 7: full GPU suite with the new parity tests (half reg, NaN, vendored reference on CUDA, config kwargs)
python -m pytest tests -m gpu -q > gpurun_out/r02_t7.log 2>&1; tail -40 gpurun_out/r02_t7.log
