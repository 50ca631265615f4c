#!/bin/bash
set -u
mkdir -p gpurun_out
for v in sub3 sub3h; do
WARPII_B200_LIB=$PWD/warpii_b200/variants/$v.so timeout 700 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 \
      -o gpurun_out/prof_${v}_V3D3_stage -f python scripts/stage_rate.py ncu_$v V3D3 > gpurun_out/ncu_${v}_V3D3.log 2>&1
done
ls -la gpurun_out/prof_sub3*
