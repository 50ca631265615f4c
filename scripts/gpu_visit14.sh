#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/stage_rate.jsonl
timeout 600 python -m pytest tests/test_gpu_maxwell.py tests/test_gpu_sources.py -q 2>&1 | tail -2
timeout 300 python scripts/stage_rate.py mxb5 N3D 2>&1 | tail -1
for v in mxb4 mxb6 mxb8; do
WARPII_B200_LIB=$PWD/warpii_b200/variants/$v.so timeout 300 python scripts/stage_rate.py $v N3D 2>&1 | tail -1
done
