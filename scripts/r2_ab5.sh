#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --steps 60 --warmup 8 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['schedule'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== f32 WS default"; q
echo "== f64 IPL=1 TMA (old)"; ODESAT_TILE_IPL=1 ODESAT_TILE_WS=0 q --precision f64 --replicas 2048
echo "== f64 wide TMA"; ODESAT_TILE_WS=0 q --precision f64 --replicas 2048
echo "== f64 wide WS"; q --precision f64 --replicas 2048
echo "== f64 wide WS nt=640?"; ODESAT_TILE_NT=640 q --precision f64 --replicas 2048
echo "== 512 replicas"; q --replicas 512 --steps 20
echo "== 1024 replicas"; q --replicas 1024 --steps 20
echo "== full gpu tests"; timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
