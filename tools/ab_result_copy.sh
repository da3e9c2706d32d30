# A/B: decoder results as views of the call's pinned buffer (default) vs pageable copies
for v in 0 1 0 1; do
  B200DET_RESULT_COPY=$v python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('RESULT_COPY=$v', round(d['value'],1), 'img/s', round(d['ms_per_step'],4), 'ms', d['clocks']['sm_mhz'])"
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
