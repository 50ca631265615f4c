#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxwell_kernel -s 6 -c 2 \
    -o gpurun_out/prof_mx2_N3D_stage -f python scripts/stage_rate.py ncu_mx2 N3D > gpurun_out/ncu_mx2_N3D.log 2>&1
tail -1 gpurun_out/ncu_mx2_N3D.log | cut -c1-200
