python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; grep -E "passed|failed|Error" gpurun_out/pytest_gpu.log | tail -3
python bench.py --quick --steps 20 --warmup 3 --engine gather 2>&1 | tail -1 | cut -c1-200
python bench.py --quick --steps 10 --warmup 3 --workload rand50k --replicas 2048 2>&1 | tail -1 | cut -c1-200
python bench.py --quick --steps 20 --warmup 3 --engine gather --precision f64 --replicas 2048 2>&1 | tail -1 | cut -c1-200
