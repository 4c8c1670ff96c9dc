#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 900 python -m pytest tests/test_gpu_tile_ragged.py -x -q 2>&1 | tail -3
timeout 300 python scripts/ragged_probe.py --mix 2:13000,3:30000 2>&1 | tail -3 | cut -c 60-330
timeout 300 python scripts/ragged_probe.py 2>&1 | tail -3 | cut -c 80-330
