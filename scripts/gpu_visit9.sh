#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/stage_rate.jsonl
(time timeout 1500 python -m pytest tests -m gpu -q) > gpurun_out/gpu_tests_r02.log 2>&1
tail -6 gpurun_out/gpu_tests_r02.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python scripts/stage_rate.py final C2 V3D3 C4s N3D C5s 2>&1 | tail -5
WARPII_B200_LIB=$PWD/warpii_b200/variants/tma.so timeout 600 python scripts/stage_rate.py tma C2 V3D3 2>&1 | tail -2
WARPII_B200_LIB=$PWD/warpii_b200/variants/tma.so timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "one_rhs or fv_blend or two_species" 2>&1 | tail -2
bash scripts/ncu_two.sh r02f V3D3 C2
timeout 900 python tests/tools/parity_table.py r02 > gpurun_out/parity_r02.log 2>&1; tail -3 gpurun_out/parity_r02.log
