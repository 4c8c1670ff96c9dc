#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 900 python -m pytest tests/test_gpu_tile_adaptive.py tests/test_gpu_tile_ragged.py tests/test_gpu_driver.py -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
