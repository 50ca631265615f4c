"""End-to-end rate of warpii_gpu_host_ssprk2_step (state in pinned host memory, whole state over PCIe both ways every step)
against the plain upload / step / download sequence, for several slab counts.
Usage: python scripts/host_step_rate.py [workload=N3D] [steps=6] [slab counts ...]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from warpii_b200 import BoxSolver, lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "N3D"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
slabs = [int(v) for v in sys.argv[3:]] or [8, 16, 32, 64, 128]
w = bench.WORKLOADS[name]
g = BoxSolver(w["dim"], w["p"], w["nx"], w["left"], w["right"], gamma=w["gamma"], **bench.species_kwargs(w))
if w.get("sources"):
    g.set_sources(True, **w["sources"])
if w.get("maxwell"):
    g.set_maxwell(True, **w["maxwell"])
host = torch.empty(g.n_dofs, dtype=torch.float64).pin_memory().numpy()
host[:] = bench.build_ic(w, g.node_coords()).reshape(-1)
g.upload(0, host)
dt = g.recommend_dt(0)
L = lib()
hp = host.ctypes.data_as(C.POINTER(C.c_double))


def plain(dt):
    assert L.warpii_gpu_upload_state(g.ctx, 0, hp, None) == 0
    g.ssprk2_step(dt, 0.0)
    assert L.warpii_gpu_download_state(g.ctx, 0, hp, None) == 0
    return g.recommend_dt(0)


for label, fn in [("plain upload/step/download", None)] + [(f"streamed, {s} slabs", s) for s in slabs]:
    d = dt
    for _ in range(2):
        d = plain(d) if fn is None else g.host_step(host, host, d, 0.0, n_slabs=fn)
    g.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        d = plain(d) if fn is None else g.host_step(host, host, d, 0.0, n_slabs=fn)
    g.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    print(f"{name} {label:32s} {ms:8.3f} ms/step  {2 * g.n_dofs / (ms * 1e-3):.3e} DoF-updates/s  "
          f"({2 * 8 * g.n_dofs / (ms * 1e-3) / 1e9:.1f} GB/s over PCIe, both directions)", flush=True)
g.close()
