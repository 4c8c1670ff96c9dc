#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 900 python -m pytest tests/test_gpu_tile_adaptive.py -x -q 2>&1 | tail -3
echo "== default depth (3 where it fits)"
timeout 300 python scripts/adaptive_probe.py --f64 2>&1 | grep tile | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['schedule'], d['precision'], round(d['ms_per_step'], 4), 'ms/step', round(d['roofline_frac'], 3))"
echo "== ring of 2"
ODESAT_TILE_AD=2 timeout 300 python scripts/adaptive_probe.py --f64 2>&1 | grep tile | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['schedule'], d['precision'], round(d['ms_per_step'], 4), 'ms/step', round(d['roofline_frac'], 3))"
