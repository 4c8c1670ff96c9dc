#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['schedule'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
for i in 1 2; do echo "== exact 4096 steps20"; q --schedule exact --steps 20 --warmup 5; done
echo "== exact 4096 steps60"; q --schedule exact --steps 60 --warmup 8
echo "== exact 4096 steps60 queue nsub=1"; ODESAT_TILE_KSUB=64 q --schedule exact --steps 60 --warmup 8
echo "== exact 4096 steps20 queue nsub=1"; ODESAT_TILE_KSUB=64 q --schedule exact --steps 20 --warmup 5
echo "== f64 exact 2048 steps60"; q --schedule exact --steps 60 --warmup 8 --precision f64 --replicas 2048
