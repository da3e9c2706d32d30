# final N=1 evidence: profile capture, bench line (default flags), reference arm, per-config table, small-batch breakdown
bash profiles/capture.sh r01 > gpurun_out/capture.log 2>&1; tail -2 gpurun_out/capture.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 400 gpurun_out/final_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -c 200 gpurun_out/final_ref.json
python tools/bench_configs.py > gpurun_out/configs_final.json 2> gpurun_out/configs_final.err
python tools/prof_small.py > gpurun_out/small_batches.json 2> gpurun_out/small_batches.err
python tests/perf_queries.py --batch 16 > gpurun_out/queries_b16.json 2> gpurun_out/queries_b16.err; cat gpurun_out/queries_b16.json
