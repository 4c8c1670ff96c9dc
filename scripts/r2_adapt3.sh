#!/bin/bash
export ODESAT_SKIP_BUILD=1
# plain run first (must exit 0), then the capture: the BALANCED f32 adaptive kernel, a 4-step launch
timeout 300 python scripts/adaptive_probe.py --steps 4 --warmup 2 > gpurun_out/plain_adapt.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_adaptive -s 5 -c 1 -f -o gpurun_out/prof_r02_tile_adaptive \
   python scripts/adaptive_probe.py --steps 4 --warmup 2 > gpurun_out/ncu_adapt.log 2>&1
tail -3 gpurun_out/ncu_adapt.log
