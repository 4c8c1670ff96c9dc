#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --steps 60 --warmup 8 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['schedule'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== 704/4 (default)"; q
echo "== 576/5"; ODESAT_TILE_NT=576 q
echo "== 640/4"; ODESAT_TILE_NT=640 ODESAT_TILE_640R4=1 q
echo "== 640/3 ws"; ODESAT_TILE_NT=640 ODESAT_TILE_WS=1 q
echo "== 704/4 (default)"; q
