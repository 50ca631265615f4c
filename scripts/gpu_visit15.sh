#!/bin/bash
set -u
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests/test_gpu_multi.py -q -x) > gpurun_out/gpu_multi_r02.log 2>&1; tail -4 gpurun_out/gpu_multi_r02.log
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3) > gpurun_out/bench_r02_2gpu.log 2>&1; grep '^{' gpurun_out/bench_r02_2gpu.log | cut -c1-1200
