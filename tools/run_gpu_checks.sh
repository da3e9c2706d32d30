python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
python tests/perf_queries.py --batch 16 > gpurun_out/queries_b16.json 2> gpurun_out/queries_b16.err; cat gpurun_out/queries_b16.json
python tools/bench_logits.py --batch 256 > gpurun_out/logits_b256.json 2> gpurun_out/logits_b256.err; cat gpurun_out/logits_b256.json
python tools/bench_logits.py --batch 256 --dtype f16 > gpurun_out/logits_b256_f16.json 2> gpurun_out/logits_b256_f16.err; cat gpurun_out/logits_b256_f16.json
