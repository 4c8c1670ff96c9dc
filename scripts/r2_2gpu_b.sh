#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 600 python -m pytest tests/test_gpu_tile_adaptive.py tests/test_gpu_driver.py tests/test_gpu_cli.py -m gpu -x -q -k "multi_gpu or devices_of_one_process or every_visible_gpu" 2>&1 | tail -2
python - <<'PY'
import sys
sys.path.insert(0, '.')
from odesat_b200 import cnf
f = cnf.random_ksat(50_000, 4.25, seed=20240611 + 4)
open('gpurun_out/rand50k.cnf', 'w').write(cnf.to_dimacs(f))
PY
for g in 1 2; do
  timeout 300 ./odesat_b200/csrc/odesat_b200_cli inter -f gpurun_out/rand50k.cnf -b 2048 -s 0.01 -n 64 --seed 1 --f32 --gpus $g --chunk 32 > gpurun_out/cli_inter_${g}gpu.out 2> gpurun_out/cli_inter_${g}gpu.err
  echo "cli gpus=$g rc=$?"; tail -n 1 gpurun_out/cli_inter_${g}gpu.err
done
cmp gpurun_out/cli_inter_1gpu.out gpurun_out/cli_inter_2gpu.out && echo "same output on 1 and 2 GPUs"
rm -f gpurun_out/rand50k.cnf
