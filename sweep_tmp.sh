python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; grep -E "passed|failed|Error" gpurun_out/pytest_gpu.log | tail -3
for s in exact balanced; do
python bench.py --quick --steps 1000 --warmup 3 --workload hard --replicas 100 --schedule $s 2>&1 | tail -1 | cut -c1-250
ODESAT_TILE_SMALL=0 python bench.py --quick --steps 1000 --warmup 3 --workload hard --replicas 100 --schedule $s 2>&1 | tail -1 | cut -c1-250
python bench.py --quick --steps 1000 --warmup 3 --workload hard --replicas 100 --schedule $s --precision f64 2>&1 | tail -1 | cut -c1-250
python bench.py --quick --steps 1000 --warmup 3 --workload hard --replicas 10000 --schedule $s 2>&1 | tail -1 | cut -c1-250
done
