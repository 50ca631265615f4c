#!/bin/bash
set -u
mkdir -p gpurun_out
(time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 3) > gpurun_out/bench_r02_8gpu.log 2>&1; grep '^{' gpurun_out/bench_r02_8gpu.log | cut -c1-600; tail -4 gpurun_out/bench_r02_8gpu.log | cut -c1-300
