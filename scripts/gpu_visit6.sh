#!/bin/bash
set -u
mkdir -p gpurun_out
for mode in mx nomx; do
  if [ $mode = nomx ]; then export WARPII_NO_MAXWELL=1; fi
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:stage_kernel -s 4 -c 4 --csv --log-file gpurun_out/n3d_$mode.csv python scripts/stage_rate.py ncu N3D > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/n3d_$mode.csv")) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    print("$mode", r[h.index("ID")], r[h.index("Metric Name")], r[h.index("Metric Value")])
PY
done
