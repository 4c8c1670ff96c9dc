#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --steps 20 --warmup 5 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['schedule'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== exact 4096"; q --schedule exact
echo "== exact 512 queue"; q --schedule exact --replicas 512
echo "== exact 512 no queue"; ODESAT_TILE_QUEUE=0 q --schedule exact --replicas 512
echo "== exact 1024 queue"; q --schedule exact --replicas 1024
echo "== exact 1024 no queue"; ODESAT_TILE_QUEUE=0 q --schedule exact --replicas 1024
echo "== balanced 4096"; q
echo "== full gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
