#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 300 python scripts/ragged_probe.py --mix 2:13000,3:30000 2>&1 | tail -3
timeout 300 python scripts/ragged_probe.py --mix 3:43000 2>&1 | tail -3
