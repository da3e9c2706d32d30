cat > /tmp/as1.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
from b200det import synth, losses
B = int(sys.argv[1])
preds = synth.make_retina_preds(B, 800, 80, seed=1, device='cuda')
ann = synth.make_annotations(B, 100, 800, 80, seed=2).cuda()
crit = losses.RetinaLoss(**synth.RETINA_KW, box_loss_type='GIoU')
with torch.no_grad():
    for _ in range(4): crit(preds, ann)
torch.cuda.synchronize()
PY
ncu --set full --clock-control none --import-source on -k regex:"retina_assign" -s 2 -c 1 -o gpurun_out/r02_assign_tile_b32 python /tmp/as1.py 32 > gpurun_out/r02_ncu_assign.log 2>&1
tail -2 gpurun_out/r02_ncu_assign.log
