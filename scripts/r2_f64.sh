#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --steps 40 --warmup 6 --precision f64 --replicas 2048 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['schedule'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== f64 default (TMA wide)"; q
echo "== f64 WS 704/4"; ODESAT_TILE_WS=1 ODESAT_TILE_NT=704 q
echo "== f64 WS 768/3"; ODESAT_TILE_WS=1 ODESAT_TILE_NT=768 q
echo "== f64 TMA 704"; ODESAT_TILE_WS=0 ODESAT_TILE_NT=704 ODESAT_TILE_TMA=1 q
echo "== f64 256 replicas default"; q --replicas 256 --steps 20
echo "== f64 256 replicas WS 704/4"; ODESAT_TILE_WS=1 ODESAT_TILE_NT=704 q --replicas 256 --steps 20
