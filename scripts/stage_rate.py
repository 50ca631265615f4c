#!/usr/bin/env python
"""Average stage-kernel time per workload for the library build selected by WARPII_B200_LIB / WARPII_GPU_STAGE (tuning A/B).
usage (GPU box): python scripts/stage_rate.py <tag> C2 V3D3 N3D ...   -> one JSON line per workload on stdout and appended to
gpurun_out/stage_rate.jsonl.  Numbers: CUDA events around every stage launch inside the library, 10 timed SSPRK2 steps
after 3 warm-up steps; not a bench value (bench.py is), only a ranking of kernel variants measured in one visit."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from warpii_b200 import BoxSolver, elems_per_block  # noqa: E402


def main():
    tag, names = sys.argv[1], sys.argv[2:]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for name in names:
        w = bench.WORKLOADS[name]
        g = BoxSolver(w["dim"], w["p"], w["nx"], w["left"], w["right"], gamma=w["gamma"], **bench.species_kwargs(w))
        if w.get("sources"):
            g.set_sources(True, **w["sources"])
        if w.get("maxwell") and not os.environ.get("WARPII_NO_MAXWELL"):
            g.set_maxwell(True, **w["maxwell"])
        u0 = bench.build_ic(w, g.node_coords())
        g.upload(0, u0)
        t, _ = g.advance_to(0.0, 1e30, max_steps=3)
        g.stage_timing(True)
        t, steps = g.advance_to(t, 1e30, max_steps=10)
        ms, n = g.stage_timing(False)
        ok = bool(np.isfinite(g.download(0)).all())
        rec = {"tag": tag, "workload": name, "stage_ms": ms / max(n, 1), "launches": n, "finite": ok,
               "patch": elems_per_block(w["dim"], w["p"]), "n_dofs": g.n_dofs,
               "hbm_frac": 20.0 * g.n_dofs / (ms / max(n, 1) * 1e-3) / 1e9 / bench.measured_peaks()[0]}
        print(json.dumps(rec), flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "stage_rate.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
        g.close()


if __name__ == "__main__":
    main()
