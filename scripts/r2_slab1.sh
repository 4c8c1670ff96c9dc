#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --workload rand50k --replicas 2048 --steps 16 --warmup 4 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['engine'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== slab tests"; timeout 1500 python -m pytest tests/test_gpu_slab.py -m gpu -x -q 2>&1 | tail -5
echo "== gather 50k"; ODESAT_SLAB=0 q --engine gather
echo "== slab var0 (occ2, pipe)"; q
echo "== slab f64 1024"; q --precision f64 --replicas 1024
echo "== gather f64 1024"; ODESAT_SLAB=0 q --precision f64 --replicas 1024 --engine gather
