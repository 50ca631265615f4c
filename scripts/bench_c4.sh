#!/bin/bash
# BASELINE config 4 at its stated size: 128^3 elements, degree 4, on 8 GPUs (under gpurun --gpus 8)
set -u
mkdir -p gpurun_out
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --workload C4 --no-other-workloads) > gpurun_out/bench_r02_C4_8gpu.log 2>&1
grep '^{' gpurun_out/bench_r02_C4_8gpu.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['config']['n_dofs'], d['value'], d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d.get('parity_check',{}).get('pass'))"
tail -3 gpurun_out/bench_r02_C4_8gpu.log | cut -c1-200
