#!/bin/bash
# final 8-/4-GPU bench lines (strong scaling of configs[2], `inter` at configs[4], adaptive steps in the same run) and the
# in-process multi-GPU tests
export ODESAT_SKIP_BUILD=1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r02h_8gpu.json 2> gpurun_out/bench_r02h_8gpu.err
echo "bench8 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 5 --inter-chunks 0 > gpurun_out/bench_r02h_4gpu.json 2> gpurun_out/bench_r02h_4gpu.err
echo "bench4 rc=$?"
timeout 600 python -m pytest tests/test_gpu_tile_adaptive.py tests/test_gpu_driver.py tests/test_gpu_cli.py -m gpu -x -q -k "multi_gpu or devices_of_one_process or every_visible_gpu" 2>&1 | tail -2
