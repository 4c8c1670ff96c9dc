#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 300 python scripts/ragged_probe.py --steps 4 --warmup 2 > gpurun_out/plain_ragged.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_ragged -s 2 -c 1 -f -o gpurun_out/prof_r02_tile_ragged \
   python scripts/ragged_probe.py --steps 4 --warmup 2 > gpurun_out/ncu_ragged.log 2>&1
tail -3 gpurun_out/ncu_ragged.log
