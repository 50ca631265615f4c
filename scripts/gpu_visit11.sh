#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/stage_rate.jsonl
timeout 300 python scripts/stage_rate.py base C3 C4s 2>&1 | tail -2
WARPII_B200_LIB=$PWD/warpii_b200/variants/mb2.so timeout 600 python scripts/stage_rate.py mb2 C3 C4s 2>&1 | tail -2
# the two kernels of one N3D first stage (pencil stage kernel with the field components skipped, stand-alone field kernel)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pencil_stage_kernel|maxwell_kernel' -s 12 -c 2 \
    -o gpurun_out/prof_r02f_N3D_stage -f python scripts/stage_rate.py ncu_r02f N3D > gpurun_out/ncu_r02f_N3D.log 2>&1
tail -2 gpurun_out/ncu_r02f_N3D.log | cut -c1-200
# the same two kernels of a second stage
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pencil_stage_kernel|maxwell_kernel' -s 14 -c 2 \
    -o gpurun_out/prof_r02f_N3D_stage2 -f python scripts/stage_rate.py ncu_r02f N3D > gpurun_out/ncu_r02f_N3D2.log 2>&1
# launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launch_list.csv \
    python bench.py --steps 2 --warmup 1 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-300
# the bench itself, both arms
(time timeout 1200 python bench.py --impl reference) > gpurun_out/bench_r02_reference.log 2>&1; tail -4 gpurun_out/bench_r02_reference.log | cut -c1-600
(time timeout 1200 python bench.py) > gpurun_out/bench_r02_default.log 2>&1; tail -4 gpurun_out/bench_r02_default.log | cut -c1-3000
