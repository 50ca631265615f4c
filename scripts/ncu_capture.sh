#!/bin/bash
# One full ncu capture of the stage kernel (2nd stage of a warm step) of the default bench workload.
# usage (on the GPU box): bash scripts/ncu_capture.sh <tag> [bench args]
set -u
tag=${1:-vX}; shift || true
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:stage_kernel -s 6 -c 1 \
    -o gpurun_out/prof_${tag}_stage -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
