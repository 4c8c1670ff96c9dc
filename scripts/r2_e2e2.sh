#!/bin/bash
export ODESAT_SKIP_BUILD=1
for i in 1 2; do
echo "== early results on";  python scripts/e2e_probe.py 20 4096 2>/dev/null | grep '"sub_batches": 8' | cut -c 1-140
echo "== early results off"; ODESAT_EARLY_RESULTS=0 python scripts/e2e_probe.py 20 4096 2>/dev/null | grep '"sub_batches": 8' | cut -c 1-140
done
echo "== 512 replicas"; python scripts/e2e_probe.py 20 512 2>/dev/null | grep '"sub_batches": 1,' | cut -c 1-140
ODESAT_EARLY_RESULTS=0 python scripts/e2e_probe.py 20 512 2>/dev/null | grep '"sub_batches": 1,' | cut -c 1-140
timeout 600 python -m pytest tests/test_gpu_driver.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
