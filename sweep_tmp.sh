python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
python bench.py --quick --steps 64 --warmup 3 --schedule exact 2>&1 | tail -1 | cut -c1-150
python bench.py --quick --steps 64 --warmup 3 2>&1 | tail -1 | cut -c1-150
