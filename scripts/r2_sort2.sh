#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "sorted_clause_view or large_single" 2>&1 | tail -3
ODESAT_GATHER_SORT=1 ODESAT_SMALL=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_commands.py tests/test_gpu_cli.py -x -q 2>&1 | tail -3
