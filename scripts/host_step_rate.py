"""End-to-end rate of warpii_gpu_host_ssprk2_step (state in pinned host memory, whole state over PCIe both ways every step)
against the plain upload / step / download sequence, for several slab counts.  C2 workload.
Usage: python scripts/host_step_rate.py [n] [steps]"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dgsem_cases as cases
from warpii_b200 import BoxSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
g = BoxSolver(2, 3, [n, n], [0.0, -5.0], [10.0, 5.0], gamma=1.4)
host = torch.empty(g.n_dofs, dtype=torch.float64).pin_memory().numpy()
host[:] = cases.to_state(cases.isentropic_vortex(1.4)(g.node_coords()), 1.4).reshape(-1)
g.upload(0, host)
dt = g.recommend_dt(0)

import ctypes as C
from warpii_b200 import lib
L = lib()
hp = host.ctypes.data_as(C.POINTER(C.c_double))

def plain(dt):
    assert L.warpii_gpu_upload_state(g.ctx, 0, hp, None) == 0
    g.ssprk2_step(dt, 0.0)
    assert L.warpii_gpu_download_state(g.ctx, 0, hp, None) == 0
    return g.recommend_dt(0)

for label, fn in [("plain upload/step/download", None)] + [(f"streamed, {s} slabs", s) for s in (4, 8, 16, 32, 64, 128)]:
    d = dt
    for _ in range(2):
        d = plain(d) if fn is None else g.host_step(host, host, d, 0.0, n_slabs=fn)
    g.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        d = plain(d) if fn is None else g.host_step(host, host, d, 0.0, n_slabs=fn)
    g.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    print(f"{label:32s} {ms:7.3f} ms/step  {2 * g.n_dofs / (ms * 1e-3):.3e} DoF-updates/s  ({2 * 8 * g.n_dofs / (ms * 1e-3) / 1e9:.1f} GB/s over PCIe, both directions)", flush=True)
g.close()
