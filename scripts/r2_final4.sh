#!/bin/bash
export ODESAT_SKIP_BUILD=1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r02e_1gpu.json 2> gpurun_out/bench_r02e_1gpu.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r02e_1gpu.json
(python scripts/adaptive_probe.py --f64; python scripts/ragged_probe.py; python scripts/ragged_probe.py --mix 2:13000,3:30000) > gpurun_out/r02_adaptive_and_ragged.jsonl 2> gpurun_out/probe.err
wc -l gpurun_out/r02_adaptive_and_ragged.jsonl
