cat > /tmp/dec1.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from b200det import synth, decode
B = int(sys.argv[1])
preds = synth.make_retina_preds(B, 800, 80, seed=1, device='cuda')
dec = decode.RetinaDecoder(**synth.RETINA_KW)
for _ in range(6): dec(preds)
torch.cuda.synchronize()
PY
for B in 1 16; do
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"select|score_argmax" -s 12 -c 4 --csv python /tmp/dec1.py $B 2>/dev/null | grep -v "^==" | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print('B=$B', d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
"
done
