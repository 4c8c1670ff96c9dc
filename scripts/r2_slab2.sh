#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --workload rand50k --replicas 2048 --steps 16 --warmup 4 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['engine'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== slab tests"; timeout 1500 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -3
echo "== lag1 bufs3 p8 (default)"; q
echo "== lag2 bufs4 p8"; ODESAT_SLAB_LAG=2 ODESAT_SLAB_BUFS=4 q
echo "== lag2 bufs4 p4"; ODESAT_SLAB_LAG=2 ODESAT_SLAB_BUFS=4 ODESAT_SLAB_CPASSES=4 ODESAT_SLAB_VPASSES=4 q
echo "== lag1 bufs3 p2"; ODESAT_SLAB_CPASSES=2 ODESAT_SLAB_VPASSES=2 q
echo "== lag1 bufs3 p4"; ODESAT_SLAB_CPASSES=4 ODESAT_SLAB_VPASSES=4 q
echo "== lag2 bufs5 p4"; ODESAT_SLAB_LAG=2 ODESAT_SLAB_BUFS=5 ODESAT_SLAB_CPASSES=4 ODESAT_SLAB_VPASSES=4 q
echo "== lag3 bufs5 p8"; ODESAT_SLAB_LAG=3 ODESAT_SLAB_BUFS=5 q
