#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/stage_rate.jsonl
timeout 300 python scripts/stage_rate.py base C2 V3D3 2>&1 | tail -2
for v in sub3 sub3h; do
  WARPII_B200_LIB=$PWD/warpii_b200/variants/$v.so timeout 600 python scripts/stage_rate.py $v C2 V3D3 N3D 2>&1 | tail -3
  WARPII_B200_LIB=$PWD/warpii_b200/variants/$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_maxwell.py -q -x 2>&1 | tail -2
done
python - <<'PY'
import sys; sys.path.insert(0,'tests')
from test_gpu_point_physics import point_fluxes
g=5/3
left = [1.0, 1.0, 0.0, 0.0, 0.5 * 1.0 + 1.0 / (g - 1.0)]
right = [0.5, 0.5, 0.0, 0.0, 0.5 * 0.5 + 0.5 / (g - 1.0)]
ec,_,_ = point_fluxes([left],[right],0,g)
print("EC golden case:", [float(x).hex() for x in ec[0]], [repr(float(x)) for x in ec[0]])
PY
