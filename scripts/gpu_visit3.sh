#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/stage_rate.jsonl
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests_r02.log 2>&1
tail -6 gpurun_out/gpu_tests_r02.log
timeout 600 python scripts/stage_rate.py pencil C2 V3D3 N3D C4s C5s 2>&1 | tail -6
WARPII_NO_MAXWELL=1 timeout 600 python scripts/stage_rate.py pencil_nomaxwell N3D C5s 2>&1 | tail -3
WARPII_GPU_STAGE=node timeout 600 python scripts/stage_rate.py node C2 V3D3 N3D C4s C5s 2>&1 | tail -6
timeout 900 python tests/tools/parity_table.py r02 > gpurun_out/parity_r02.log 2>&1; tail -20 gpurun_out/parity_r02.log
(time timeout 900 python bench.py) > gpurun_out/bench_default_r02.log 2>&1; tail -c 3000 gpurun_out/bench_default_r02.log
