#!/bin/bash
export ODESAT_SKIP_BUILD=1
echo "== parity tests with the sorted view forced on every single instance"
ODESAT_GATHER_SORT=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_commands.py -x -q 2>&1 | tail -3
echo "== parity tests, default"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "large_single or golden or simulate" 2>&1 | tail -2
q() { python bench.py --quick --workload rand1m --replicas 1 --steps 60 --warmup 10 "$@" 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],5), 'ms/step  frac', round(d['frac'],4), d['engine'], d['precision'])"; }
for i in 1 2; do
echo "== rand1m sorted (default)"; q
echo "== rand1m unsorted"; ODESAT_GATHER_SORT=0 q
done
echo "== f64 sorted / unsorted"; q --precision f64; ODESAT_GATHER_SORT=0 q --precision f64
