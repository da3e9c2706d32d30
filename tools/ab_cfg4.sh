# A/B of the raw-tile arg-max sweep knobs at BASELINE configs[3]
for env in "" "B200DET_RAW_NO_BULK=1"; do
  echo "[$env] $(env $env python tools/prof_cfg4.py 2>&1 | tail -1)"
done
