#!/bin/bash
# A/B: warp-specialised persistent kernel + wide BALANCED levels vs the TMA kernel / per-thread ring
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --steps 60 --warmup 8 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['schedule'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== parity"; timeout 900 python -m pytest tests/test_gpu_tile.py -m gpu -x -q -k "warp_specialised" 2>&1 | tail -5
echo "== old default (IPL=1, TMA)";           ODESAT_TILE_IPL=1 ODESAT_TILE_WS=0 q
echo "== IPL=1 WS";                           ODESAT_TILE_IPL=1 ODESAT_TILE_WS=1 q
echo "== wide, TMA kernel (barrier per item)"; ODESAT_TILE_WS=0 q
echo "== wide, ring kernel";                  ODESAT_TILE_WS=0 ODESAT_TILE_TMA=0 q
echo "== wide, WS (new default)";             q
echo "== wide, WS D=2";                       ODESAT_TILE_D=2 q
echo "== wide, WS ksub=10";                   ODESAT_TILE_KSUB=10 q
echo "== wide, WS ksub=4";                    ODESAT_TILE_KSUB=4 q
echo "== wide, WS nt=512";                    ODESAT_TILE_NT=512 ODESAT_TILE_WS=1 q
echo "== exact default";                      q --schedule exact
echo "== exact WS nt=512";                    ODESAT_TILE_WS=1 q --schedule exact
echo "== f64 old";                            ODESAT_TILE_IPL=1 ODESAT_TILE_WS=0 q --precision f64 --replicas 2048
echo "== f64 wide WS";                        q --precision f64 --replicas 2048
echo "== 512 replicas (8-GPU share) old";     ODESAT_TILE_IPL=1 ODESAT_TILE_WS=0 q --replicas 512 --steps 20
echo "== 512 replicas wide WS";               q --replicas 512 --steps 20
echo "== 1024 replicas (4-GPU share) old";    ODESAT_TILE_IPL=1 ODESAT_TILE_WS=0 q --replicas 1024 --steps 20
echo "== 1024 replicas wide WS";              q --replicas 1024 --steps 20
