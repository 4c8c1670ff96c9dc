#!/bin/bash
export ODESAT_SKIP_BUILD=1
python bench.py --steps 16 --warmup 3 --inter-chunks 0 > gpurun_out/b_launch.log 2>&1 || { tail -5 gpurun_out/b_launch.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02e_launches_bench_steps16.csv \
    python bench.py --steps 16 --warmup 3 --inter-chunks 0 > gpurun_out/ncu_launch.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/r02e_launches_bench_steps16.csv
