#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02i_1gpu.json 2> gpurun_out/bench_r02i_1gpu.err; echo "bench rc=$?"
timeout 600 python scripts/run_configs.py > gpurun_out/r02i_configs.jsonl 2> gpurun_out/r02i_configs.err; echo "configs rc=$?"
