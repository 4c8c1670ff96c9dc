python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; grep -E "passed|failed|Error" gpurun_out/pytest_gpu.log | tail -3
python bench.py --quick --steps 64 --warmup 3 2>&1 | tail -1 | cut -c1-200
python bench.py --quick --steps 64 --warmup 3 --schedule exact 2>&1 | tail -1 | cut -c1-200
ODESAT_TILE_NT=768 python bench.py --quick --steps 64 --warmup 3 --schedule exact 2>&1 | tail -1 | cut -c1-200
python bench.py --quick --steps 64 --warmup 3 --precision f64 --replicas 2048 --schedule exact 2>&1 | tail -1 | cut -c1-200
python bench.py --quick --steps 64 --warmup 3 --precision f64 --replicas 2048 2>&1 | tail -1 | cut -c1-200
