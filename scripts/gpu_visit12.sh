#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/stage_rate.jsonl
timeout 600 python -m pytest tests/test_gpu_maxwell.py tests/test_gpu_sources.py -q 2>&1 | tail -3
timeout 300 python scripts/stage_rate.py mx2 N3D 2>&1 | tail -1
WARPII_GPU_STAGE=node timeout 300 python scripts/stage_rate.py mx2_node C5s 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:maxwell_kernel -s 6 -c 2 python scripts/stage_rate.py ncu_mx2 N3D 2>&1 | grep -E "maxwell_kernel|duration|dram__" | head -8
