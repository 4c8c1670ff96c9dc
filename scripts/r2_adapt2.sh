#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 900 python -m pytest tests/test_gpu_tile_adaptive.py -x -q 2>&1 | tail -5
timeout 300 python scripts/adaptive_probe.py 2>&1 | tail -3
echo "== e2e 512 replicas, 20 steps"
timeout 300 python scripts/e2e_probe.py 20 512 2>&1 | tail -5
