#!/bin/bash
# A/B: shared-memory ring (TMA / cp.async) vs L2-prefetched register pipeline
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --steps 64 --warmup 8 "$@" 2>/dev/null | tail -1; }
echo "== default balanced"; q
echo "== default exact"; q --schedule exact
for d in 0 2 4; do
  echo "== pf balanced dist=$d"; ODESAT_TILE_PIPE=pf ODESAT_TILE_PF_DIST=$d q
done
for nt in 512 640 1024; do
  echo "== pf balanced nt=$nt"; ODESAT_TILE_PIPE=pf ODESAT_TILE_NT=$nt q
done
echo "== pf exact"; ODESAT_TILE_PIPE=pf q --schedule exact
echo "== pf exact nt=768"; ODESAT_TILE_PIPE=pf ODESAT_TILE_NT=768 q --schedule exact
echo "== pf f64 balanced"; ODESAT_TILE_PIPE=pf q --precision f64 --replicas 2048
echo "== parity with pf"
ODESAT_TILE_PIPE=pf timeout 600 python -m pytest tests/test_gpu_tile.py -m gpu -x -q 2>&1 | tail -3
