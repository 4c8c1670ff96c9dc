#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --steps 60 --warmup 8 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['schedule'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== pf 0 (default)"; q
for p in 1 2 4 8 16; do echo "== pf $p"; ODESAT_TILE_WS_PF=$p q; done
echo "== pf 0 (default)"; q
echo "== 768/3 pf 4"; ODESAT_TILE_NT=768 ODESAT_TILE_WS_PF=4 q
echo "== 768/3 pf 0"; ODESAT_TILE_NT=768 q
