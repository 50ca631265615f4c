#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 700 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 7 -c 1 -o gpurun_out/prof_p3_N3D_stage2 -f python scripts/stage_rate.py ncu N3D > gpurun_out/ncu_p3.log 2>&1
tail -2 gpurun_out/ncu_p3.log | cut -c1-200
