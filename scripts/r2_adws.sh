#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 300 python -m pytest tests/test_gpu_tile_adaptive.py -x -q -k "warp_specialised" 2>&1 | tail -5
echo "== ws on";  timeout 120 python scripts/adaptive_probe.py 2>&1 | grep balanced | cut -c 60-260
echo "== ws off"; ODESAT_TILE_ADWS=0 timeout 120 python scripts/adaptive_probe.py 2>&1 | grep balanced | cut -c 60-260
echo "== 512 replicas ws on";  timeout 120 python scripts/adaptive_probe.py --replicas 512 2>&1 | grep balanced | cut -c 60-260
echo "== 512 replicas ws off"; ODESAT_TILE_ADWS=0 timeout 120 python scripts/adaptive_probe.py --replicas 512 2>&1 | grep balanced | cut -c 60-260
