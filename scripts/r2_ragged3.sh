#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 900 python -m pytest tests/test_gpu_tile_ragged.py -x -q 2>&1 | tail -12
echo "== groups on"
timeout 300 python scripts/ragged_probe.py 2>&1 | tail -3
echo "== groups off"
ODESAT_TILE_GROUPS=0 timeout 300 python scripts/ragged_probe.py 2>&1 | tail -3 | head -2
