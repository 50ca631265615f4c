#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/stage_rate.jsonl
WL="C2 V3D3 C4s"
timeout 600 python scripts/stage_rate.py base $WL 2>&1 | tail -3
WARPII_GPU_MAXWELL=fused timeout 600 python scripts/stage_rate.py base_fused N3D C5s 2>&1 | tail -2
WARPII_GPU_MAXWELL=separate timeout 600 python scripts/stage_rate.py base_sep N3D C5s 2>&1 | tail -2
for v in p0roll rt2 rt2roll; do
  WARPII_B200_LIB=$PWD/warpii_b200/variants/$v.so timeout 600 python scripts/stage_rate.py $v $WL 2>&1 | tail -3
  WARPII_GPU_MAXWELL=fused WARPII_B200_LIB=$PWD/warpii_b200/variants/$v.so timeout 600 python scripts/stage_rate.py ${v}_fused N3D C5s 2>&1 | tail -2
  WARPII_GPU_MAXWELL=separate WARPII_B200_LIB=$PWD/warpii_b200/variants/$v.so timeout 600 python scripts/stage_rate.py ${v}_sep N3D 2>&1 | tail -1
done
