#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/stage_rate.jsonl
timeout 300 python scripts/stage_rate.py base C2 V3D3 N3D 2>&1 | tail -3
for v in pf1 pf2 p0roll rtdir2; do
  WARPII_B200_LIB=$PWD/warpii_b200/variants/$v.so timeout 600 python scripts/stage_rate.py $v C2 V3D3 N3D 2>&1 | tail -3
done
WARPII_B200_LIB=$PWD/warpii_b200/variants/pf1.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x 2>&1 | tail -1
