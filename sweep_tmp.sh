python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for sched in exact balanced; do
 for d in 2 4; do
 ODESAT_TILE_NT=1024 ODESAT_TILE_D=$d python bench.py --quick --steps 64 --warmup 3 --schedule $sched 2>&1 | tail -1
 done
 for d in 3 6; do
  ODESAT_TILE_NT=512 ODESAT_TILE_D=$d python bench.py --quick --steps 64 --warmup 3 --schedule $sched 2>&1 | tail -1
 done
done
python bench.py --quick --steps 64 --warmup 3 --precision f64 --replicas 2048 2>&1 | tail -1
python bench.py --quick --steps 64 --warmup 3 --precision f64 --replicas 2048 --schedule balanced 2>&1 | tail -1
