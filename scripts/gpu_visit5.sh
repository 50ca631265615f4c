#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/stage_rate.jsonl
(time timeout 1500 python -m pytest tests -m gpu -q -x) > gpurun_out/gpu_tests_r02.log 2>&1
tail -5 gpurun_out/gpu_tests_r02.log
timeout 600 python scripts/stage_rate.py rtdir C2 V3D3 N3D C4s C5s 2>&1 | tail -6
WARPII_NO_MAXWELL=1 timeout 600 python scripts/stage_rate.py rtdir_nomx N3D C5s 2>&1 | tail -3
WARPII_B200_LIB=$PWD/warpii_b200/variants/tdir.so timeout 600 python scripts/stage_rate.py tdir C2 V3D3 N3D C4s C5s 2>&1 | tail -6
