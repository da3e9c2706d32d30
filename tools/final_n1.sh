# final N=1 evidence: bench line (default flags), reference arm, training-sweep ncu, per-config table
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; tail -c 300 gpurun_out/final_ref.json
bash tools/ncu_train.sh gpurun_out/train_sweep2
python tools/bench_configs.py > gpurun_out/configs_final.json 2> gpurun_out/configs_final.err; tail -c 300 gpurun_out/configs_final.json
