#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02g_1gpu.json 2> gpurun_out/bench_r02g_1gpu.err; echo "bench rc=$?"
(python scripts/adaptive_probe.py --f64; python scripts/ragged_probe.py; python scripts/ragged_probe.py --mix 2:13000,3:30000; python scripts/adaptive_probe.py --mix 2:13000,3:30000) > gpurun_out/r02g_adaptive_and_ragged.jsonl 2> gpurun_out/probe.err
wc -l gpurun_out/r02g_adaptive_and_ragged.jsonl
