#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['schedule'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== exact 4096 steps20"; q --schedule exact --steps 20 --warmup 5
echo "== exact 512 queue"; q --schedule exact --replicas 512 --steps 20 --warmup 5
echo "== exact 512 no queue"; ODESAT_TILE_QUEUE=0 q --schedule exact --replicas 512 --steps 20 --warmup 5
echo "== f64 exact 2048"; q --schedule exact --steps 20 --warmup 5 --precision f64 --replicas 2048
echo "== f64 exact 256 queue"; q --schedule exact --steps 20 --warmup 5 --precision f64 --replicas 256
echo "== f64 exact 256 no queue"; ODESAT_TILE_QUEUE=0 q --schedule exact --steps 20 --warmup 5 --precision f64 --replicas 256
echo "== tile tests"; timeout 2400 python -m pytest tests/test_gpu_tile.py tests/test_gpu_driver.py -m gpu -x -q 2>&1 | tail -3
