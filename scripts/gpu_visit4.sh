#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/stage_rate.jsonl
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/gpu_tests_r02.log 2>&1
tail -12 gpurun_out/gpu_tests_r02.log
timeout 600 python scripts/stage_rate.py pencil C2 V3D3 N3D C4s C5s 2>&1 | tail -6
WARPII_GPU_STAGE=node timeout 600 python scripts/stage_rate.py node N3D C5s 2>&1 | tail -3
