# r02 (third session): criterion-only calls at small batches, 32- vs 24-location assignment tiles
for i in 1 2; do
echo "tiles 32:"; B200DET_TILE_SMALL_BATCH=0 timeout 200 python tools/prof_loss_small.py
echo "tiles 24 (batch <= 16):"; timeout 200 python tools/prof_loss_small.py
done
