#!/bin/bash
# Two full ncu captures of the stage kernel in one visit: scripts/ncu_two.sh <tag> <workloadA> <workloadB>
set -u
tag=$1; A=$2; B=${3:-}
mkdir -p gpurun_out
for w in $A $B; do
  timeout 700 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 \
      -o gpurun_out/prof_${tag}_${w}_stage -f python scripts/stage_rate.py ncu_$tag $w > gpurun_out/ncu_${tag}_${w}.log 2>&1
  tail -2 gpurun_out/ncu_${tag}_${w}.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
