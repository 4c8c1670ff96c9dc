#!/bin/bash
# 8-GPU record of round 2: the driver's bench line (strong scaling of configs[2], `inter` at configs[4] with the MIN all-reduce
# inside the timed region), the in-process multi-GPU tests, and the C++ CLI driving 8 GPUs from one process.
export ODESAT_SKIP_BUILD=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02c_8gpu.json 2> gpurun_out/bench_r02c_8gpu.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r02c_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --inter-chunks 0 > gpurun_out/bench_r02c_4gpu.json 2> gpurun_out/bench_r02c_4gpu.err
echo "bench4 rc=$?"
timeout 900 python -m pytest tests/test_gpu_driver.py tests/test_gpu_cli.py -m gpu -x -q -k "multi_gpu or every_visible_gpu" 2>&1 | tail -3
python - <<'PY'
import sys
sys.path.insert(0, '.')
from odesat_b200 import cnf
f = cnf.random_ksat(50_000, 4.25, seed=20240611 + 4)
open('gpurun_out/rand50k.cnf', 'w').write(cnf.to_dimacs(f))
PY
for g in 1 8; do
  ./odesat_b200/csrc/odesat_b200_cli inter -f gpurun_out/rand50k.cnf -b 16384 -s 0.01 -n 96 --seed 1 --f32 --gpus $g --chunk 32 > gpurun_out/cli_inter_${g}gpu.out 2> gpurun_out/cli_inter_${g}gpu.err
  echo "cli gpus=$g rc=$?"; tail -n 2 gpurun_out/cli_inter_${g}gpu.err
done
rm -f gpurun_out/rand50k.cnf
