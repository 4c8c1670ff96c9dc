#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --workload rand50k --replicas 2048 --steps 20 --warmup 5 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['engine'], d['precision'])"; }
echo "== rand50k unsorted"; q
echo "== rand50k sorted"; ODESAT_GATHER_SORT_BATCH_N=1000 q
echo "== rand50k unsorted"; q
echo "== rand50k sorted"; ODESAT_GATHER_SORT_BATCH_N=1000 q
ODESAT_GATHER_SORT=1 ODESAT_SMALL=0 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_driver.py -x -q 2>&1 | tail -3
