#!/bin/bash
# bench.py on N GPUs of this box the way the driver launches it: scripts/bench_n.sh N  (under gpurun --gpus N)
set -u
N=${1:-2}
mkdir -p gpurun_out
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N) > gpurun_out/bench_r02_${N}gpu.log 2>&1
grep '^{' gpurun_out/bench_r02_${N}gpu.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['steps'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d.get('parity_check',{}).get('pass'), d['details'].get('host_numa_binding_rank0'))"
tail -3 gpurun_out/bench_r02_${N}gpu.log | cut -c1-200
