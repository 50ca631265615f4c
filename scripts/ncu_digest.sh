#!/bin/bash
# Turn gpurun_out/prof_<tag>_stage.ncu-rep into the three text summaries kept under profiles/.
# usage (here, no GPU needed): bash scripts/ncu_digest.sh <tag> <round>
set -eu
tag=$1; rnd=${2:-r01}
cd gpurun_out
ncu -i prof_${tag}_stage.ncu-rep --page raw --csv > ${tag}_raw.csv
ncu -i prof_${tag}_stage.ncu-rep --page source --csv --print-source cuda,sass > ${tag}_cs.csv
ncu -i prof_${tag}_stage.ncu-rep --page source --csv --print-source sass > ${tag}_sass.csv
cd ..
python scripts/ncu_summary.py gpurun_out/prof_${tag}_stage.ncu-rep > profiles/${rnd}_${tag}_stage_kernel_summary.txt
python scripts/ncu_lines.py gpurun_out/${tag}_cs.csv > profiles/${rnd}_${tag}_stage_kernel_lines.txt
python scripts/ncu_opcodes.py gpurun_out/${tag}_sass.csv > profiles/${rnd}_${tag}_stage_kernel_opcodes.txt
head -30 profiles/${rnd}_${tag}_stage_kernel_summary.txt
