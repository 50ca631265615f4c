#!/bin/bash
# A/B bench of library builds (tuning experiments): default lib and every warpii_b200/variants/*.so
python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('default', d['value'], d['roofline']['avg_launch_ms'])"
for v in warpii_b200/variants/*.so; do
  WARPII_B200_LIB=$PWD/$v python bench.py --steps 20 --warmup 3 --no-cpu-baseline ${BENCH_ARGS:-} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['roofline']['avg_launch_ms'])"
done
