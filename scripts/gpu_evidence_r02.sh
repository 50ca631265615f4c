#!/bin/bash
# Round-2 evidence visit (one B200): GPU tests, smoke, stage times of every workload, ncu capture of the four N3D launches,
# launch list of the bench command, both bench arms, parity table.  Run through gpurun; outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
(free -g; nproc; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv) > gpurun_out/box.txt 2>&1
rm -f gpurun_out/stage_rate.jsonl
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/gpu_tests_r02.log 2>&1
tail -4 gpurun_out/gpu_tests_r02.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 600 python scripts/stage_rate.py final C2 C3 V3D3 C4s N3D C5s 2>&1 | tail -6
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pencil_stage_kernel|maxwell_kernel' -s 12 -c 4 \
    -o gpurun_out/prof_r02g_N3D_stage -f python scripts/stage_rate.py ncu_r02g N3D > gpurun_out/ncu_r02g_N3D.log 2>&1
tail -1 gpurun_out/ncu_r02g_N3D.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_final_launch_list.csv \
    python bench.py --steps 2 --warmup 1 --no-other-workloads --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
(time timeout 1200 python bench.py --impl reference) > gpurun_out/bench_r02_reference.log 2>&1; grep '^{' gpurun_out/bench_r02_reference.log | cut -c1-300
(time timeout 1200 python bench.py) > gpurun_out/bench_r02_default.log 2>&1; grep '^{' gpurun_out/bench_r02_default.log | cut -c1-400
timeout 900 python tests/tools/parity_table.py r02 > gpurun_out/parity_r02.log 2>&1; tail -2 gpurun_out/parity_r02.log
