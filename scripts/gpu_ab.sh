#!/bin/bash
# GPU visit: parity tests with the default build, then stage-kernel times of the default build, the node-per-thread kernel
# and every variant build under warpii_b200/variants/.  usage: bash scripts/gpu_ab.sh "<workloads>" [skip_tests]
set -u
WL=${1:-"C2 V3D3 N3D C4s"}
mkdir -p gpurun_out
rm -f gpurun_out/stage_rate.jsonl
if [ -z "${2:-}" ]; then
  (time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/gpu_tests_r02.log 2>&1
  tail -5 gpurun_out/gpu_tests_r02.log
fi
timeout 600 python scripts/stage_rate.py pencil $WL 2>&1 | tail -8
WARPII_GPU_STAGE=node timeout 600 python scripts/stage_rate.py node $WL 2>&1 | tail -8
for v in warpii_b200/variants/*.so; do
  [ -f "$v" ] || continue
  WARPII_B200_LIB=$PWD/$v timeout 600 python scripts/stage_rate.py $(basename $v .so) $WL 2>&1 | tail -8
done
