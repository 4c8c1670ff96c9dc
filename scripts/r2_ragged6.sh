#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 900 python -m pytest tests/test_gpu_tile_ragged.py tests/test_gpu_tile_adaptive.py -x -q 2>&1 | tail -12
