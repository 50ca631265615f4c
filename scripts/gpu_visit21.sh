#!/bin/bash
set -u
mkdir -p gpurun_out
(nproc; lscpu | grep -i "numa\|socket\|model name"; nvidia-smi topo -m; python -c "import os; print(sorted(os.sched_getaffinity(0)))") > gpurun_out/topology.txt 2>&1
N=${1:-2}
(time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3) > gpurun_out/bench_r02_${N}gpu_numa.log 2>&1
grep '^{' gpurun_out/bench_r02_${N}gpu_numa.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['details'].get('host_numa_binding_rank0'))"
tail -3 gpurun_out/bench_r02_${N}gpu_numa.log | cut -c1-200
