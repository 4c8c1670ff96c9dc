#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 1200 python scripts/run_configs.py > gpurun_out/r02f_configs.jsonl 2> gpurun_out/r02f_configs.err; echo "rc=$?"
wc -l gpurun_out/r02f_configs.jsonl; tail -3 gpurun_out/r02f_configs.err
