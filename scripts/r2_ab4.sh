#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --steps 60 --warmup 8 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['schedule'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_tile.py -m gpu -x -q -k "warp_specialised" 2>&1 | tail -5
echo "== wide, TMA kernel";  ODESAT_TILE_WS=0 q
for v in 0 1 2 3; do echo "== wide, WS var=$v"; ODESAT_TILE_WSVAR=$v q; done
echo "== parity var=3"; ODESAT_TILE_WSVAR=3 timeout 900 python -m pytest tests/test_gpu_tile.py -m gpu -x -q -k "warp_specialised and F32" 2>&1 | tail -3
for v in 0 3; do echo "== 512 replicas WS var=$v"; ODESAT_TILE_WSVAR=$v q --replicas 512 --steps 20; done
echo "== 512 replicas WS ksub=20 (no split)"; ODESAT_TILE_KSUB=20 q --replicas 512 --steps 20
echo "== 512 replicas WS ksub=1"; ODESAT_TILE_KSUB=1 q --replicas 512 --steps 20
echo "== 512 replicas WS ksub=2"; ODESAT_TILE_KSUB=2 q --replicas 512 --steps 20
