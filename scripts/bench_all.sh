#!/bin/bash
# One JSON line per workload with the current build (profiles/<round>_bench_lines.jsonl), plus the ncu launch list of the
# default bench command.  usage (on the GPU box): bash scripts/bench_all.sh <round>
set -u
rnd=${1:-r01}
mkdir -p gpurun_out
out=gpurun_out/${rnd}_bench_lines.jsonl
: > $out
timeout 300 python bench.py --steps 20 --warmup 3 2>&1 | tail -1 >> $out
for w in C2c C3 C3s C5s V3D3 N3D C4s; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 >> $out
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${rnd}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
python - <<PY
import json
for line in open("$out"):
    try:
        d = json.loads(line)
    except Exception:
        print("BAD LINE", line[:200]); continue
    print("%-8s value %.4g  stage %.4f ms  frac %.3f  e2e %.3g" % (d["config"]["workload"][:8], d["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"], d["e2e"]["value"]))
PY
