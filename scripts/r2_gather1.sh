#!/bin/bash
export ODESAT_SKIP_BUILD=1
q() { python bench.py --quick --workload rand50k --replicas 2048 --steps 16 --warmup 4 --engine gather "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['engine'], d['precision'], 'R', d['replicas_per_gpu'], 'launches', d['launches'])"; }
echo "== gather 50k default (packed)"; q
echo "== gather 50k scalar fast (old)"; ODESAT_GATHER_PACKED=0 q
echo "== packed + l2hints"; ODESAT_GATHER_L2HINTS=1 q
echo "== packed + bx16"; ODESAT_GATHER_BX=16 q
echo "== packed + bx16 + l2hints"; ODESAT_GATHER_BX=16 ODESAT_GATHER_L2HINTS=1 q
echo "== packed + bx32 + l2hints"; ODESAT_GATHER_BX=32 ODESAT_GATHER_L2HINTS=1 q
echo "== packed + bx8 + l2hints"; ODESAT_GATHER_BX=8 ODESAT_GATHER_L2HINTS=1 q
echo "== 10k gather"; python bench.py --quick --steps 16 --warmup 4 --engine gather 2>/dev/null | tail -1 | cut -c1-220
echo "== parity"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tile.py -m gpu -x -q 2>&1 | tail -3
