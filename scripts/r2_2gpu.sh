#!/bin/bash
export ODESAT_SKIP_BUILD=1
timeout 900 python -m pytest tests/test_gpu_tile_adaptive.py tests/test_gpu_driver.py -m gpu -x -q -k "multi_gpu or devices_of_one_process or every_visible_gpu" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02e_2gpu.json 2> gpurun_out/bench_r02e_2gpu.err
echo "bench2 rc=$?"; tail -c 600 gpurun_out/bench_r02e_2gpu.json
