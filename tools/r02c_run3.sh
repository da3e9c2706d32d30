# r02 (third session): A/B of the fused sweep's register cap (3 CTAs x 72 registers leave room for one
# 40-register assignment CTA per SM beside the sweep; 64 registers: four sweep CTAs)
# variants/lib_r72.so, lib_r64.so: decode.cu rebuilt with -DB200DET_FUSED_MAXNREG=72 / 64 and linked with the other
# objects of csrc/_obj (variants/ is scratch: built for this run, not kept)
mkdir -p gpurun_out
LIB=simpleaicv-pytorch-imagenet-coco-training_b200/libb200det.so
for v in default r72 r64; do
  cp variants/lib_$v.so $LIB
  for b in 256 32; do
    timeout 300 python bench.py --batch $b --steps 60 --warmup 5 --no-e2e --no-cpu-baseline --no-configs --no-fused --no-separate > gpurun_out/c3_${v}_b$b.json 2> gpurun_out/c3_${v}_b$b.err; echo "$v b$b rc=$?"
  done
  timeout 300 python tools/prof_eval_step.py --batch 256 2>&1 | tail -1
  B200DET_LOSS_OVERLAP=0 timeout 300 python tools/prof_eval_step.py --batch 256 2>&1 | tail -1
done
cp variants/lib_default.so $LIB
python - <<'PY'
import json
for v in ('default','r72','r64'):
    for b in (256,32):
        d=json.loads(open(f'gpurun_out/c3_{v}_b{b}.json').read().strip().splitlines()[-1])
        print(v, b, round(d['ms_per_step'],4), d['kernels_ms'], d['clocks'] and d['clocks']['sm_mhz'])
PY
