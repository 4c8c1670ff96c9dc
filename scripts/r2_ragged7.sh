#!/bin/bash
export ODESAT_SKIP_BUILD=1
python scripts/adaptive_probe.py --mix 2:13000,3:30000 2>&1 | tail -3 | cut -c 1-420
