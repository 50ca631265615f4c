#!/usr/bin/env python
"""FP64 FMA rate vs resident warps per SM and independent chains per thread (tuning diagnostic, GPU box)."""
import ctypes as C
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from warpii_b200 import lib
L = lib()
peak = C.c_double(0)
L.warpii_gpu_measure_fp64_peak(0, C.byref(peak))
out = {"peak_fma_per_s": peak.value, "rows": []}
for warps in (4, 8, 12, 16, 20, 24, 32):
    row = {"warps_per_sm": warps}
    for ilp in (1, 2, 3, 4, 6, 8):
        r = C.c_double(0)
        assert L.warpii_gpu_fp64_rate_probe(0, 32 * warps, ilp, C.byref(r)) == 0
        row[f"ilp{ilp}"] = round(r.value / peak.value, 3)
    out["rows"].append(row)
    print(row, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "fp64_probe.json"), "w"), indent=1)
