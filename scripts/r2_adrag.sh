#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 900 python -m pytest tests/test_gpu_tile_ragged.py tests/test_gpu_tile_adaptive.py -x -q 2>&1 | tail -3
timeout 300 python scripts/adaptive_probe.py --mix 2:13000,3:30000 2>&1 | grep tile | cut -c 60-300
timeout 300 python scripts/adaptive_probe.py 2>&1 | grep tile | cut -c 1-220
