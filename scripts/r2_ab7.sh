#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --steps 60 --warmup 8 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['schedule'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
for i in 1 2; do
echo "== 704 (default)"; q
echo "== 672"; ODESAT_TILE_NT=672 q
echo "== 736"; ODESAT_TILE_NT=736 q
done
echo "== parity 672"; ODESAT_TILE_NT=672 timeout 600 python -m pytest tests/test_gpu_tile.py -m gpu -x -q -k "hundred_step or balanced_f32_meets" 2>&1 | tail -2
