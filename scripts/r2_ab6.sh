#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --steps 60 --warmup 8 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['schedule'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== default"; q
echo "== early"; ODESAT_TILE_WS_EARLY=1 q
echo "== nt704 d4"; ODESAT_TILE_NT=704 q
echo "== nt704 d4 early"; ODESAT_TILE_NT=704 ODESAT_TILE_WS_EARLY=1 q
echo "== nt704 d3"; ODESAT_TILE_NT=704 ODESAT_TILE_D=3 q
echo "== parity early"; ODESAT_TILE_WS_EARLY=1 timeout 900 python -m pytest tests/test_gpu_tile.py -m gpu -x -q -k "warp_specialised" 2>&1 | tail -2
echo "== parity 704"; ODESAT_TILE_NT=704 ODESAT_TILE_WS_EARLY=1 timeout 900 python -m pytest tests/test_gpu_tile.py -m gpu -x -q -k "warp_specialised_kernel_equals" 2>&1 | tail -2
