#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --workload rand50k --replicas 2048 --steps 16 --warmup 4 --engine slab "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['engine'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== slab tests"; timeout 1500 python -m pytest tests/test_gpu_slab.py tests/test_gpu_driver.py -m gpu -x -q 2>&1 | tail -3
echo "== pipe c4 lag1 bufs3 (default)"; q
echo "== pipe c4 lag2 bufs4"; ODESAT_SLAB_LAG=2 ODESAT_SLAB_BUFS=4 q
echo "== pipe c2 lag1 bufs3"; ODESAT_SLAB_CPASSES=2 q
echo "== pipe c2 lag2 bufs4"; ODESAT_SLAB_CPASSES=2 ODESAT_SLAB_LAG=2 ODESAT_SLAB_BUFS=4 q
echo "== old warp kernel"; ODESAT_SLAB_PIPE=0 q
echo "== pipe f64 1024"; q --precision f64 --replicas 1024
