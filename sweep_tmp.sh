python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; grep -E "passed|failed|Error" gpurun_out/pytest_gpu.log | tail -3
python bench.py --workload rand1m --steps 50 --warmup 5 2>&1 | tail -1 | cut -c1-330
python bench.py --workload rand1m --steps 50 --warmup 5 --precision f64 2>&1 | tail -1 | cut -c1-330
python bench.py --quick --steps 20 --warmup 3 --engine gather 2>&1 | tail -1 | cut -c1-200
