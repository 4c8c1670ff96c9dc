#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
