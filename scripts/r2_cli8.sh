#!/bin/bash
# The C++ CLI (one process, no Python) driving 8 GPUs: `inter` at the configs[4] shape, long enough to amortise start-up.
python - <<'PY'
import sys
sys.path.insert(0, '.')
from odesat_b200 import cnf
f = cnf.random_ksat(50_000, 4.25, seed=20240611 + 4)
open('gpurun_out/rand50k.cnf', 'w').write(cnf.to_dimacs(f))
PY
for g in 8 4; do
  ./odesat_b200/csrc/odesat_b200_cli inter -f gpurun_out/rand50k.cnf -b 16384 -s 0.01 -n 1920 --seed 1 --f32 --gpus $g --chunk 32 > /dev/null 2> gpurun_out/cli_inter_long_${g}gpu.err
  echo "cli gpus=$g rc=$?"; tail -n 2 gpurun_out/cli_inter_long_${g}gpu.err
done
rm -f gpurun_out/rand50k.cnf
