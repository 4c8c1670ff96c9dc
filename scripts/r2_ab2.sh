#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --steps 64 --warmup 8 "$@" 2>/dev/null | tail -1; }
for nt in 768 1024; do for d in 6 8 12 16 24 40; do
  echo "== pf balanced nt=$nt dist=$d"; ODESAT_TILE_PIPE=pf ODESAT_TILE_NT=$nt ODESAT_TILE_PF_DIST=$d q
done; done
for nt in 512 768 1024; do for d in 8 16; do
  echo "== pf exact nt=$nt dist=$d"; ODESAT_TILE_PIPE=pf ODESAT_TILE_NT=$nt ODESAT_TILE_PF_DIST=$d q --schedule exact
done; done
echo "== pf f64 balanced nt=768 dist=12"; ODESAT_TILE_PIPE=pf ODESAT_TILE_NT=768 ODESAT_TILE_PF_DIST=12 q --precision f64 --replicas 2048
