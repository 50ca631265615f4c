"""Steps per second of launch-bound small meshes through the device-resident loop (CUDA-graph replay vs plain launches:
run once as is and once with WARPII_GPU_NO_GRAPH=1)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import dgsem_cases as cases
from warpii_b200 import BoxSolver, BC_OUTFLOW

for label, kw, ic in [("C1 1D Sod p=2 200 cells", dict(dim=1, fe_degree=2, nx=[200], left=[0.0], right=[1.0], periodic=[0], gamma=5 / 3,
                                                      bc_kinds=[[BC_OUTFLOW, BC_OUTFLOW]]), cases.sod()),
                      ("2D vortex p=3 32x32", dict(dim=2, fe_degree=3, nx=[32, 32], left=[0.0, -5.0], right=[10.0, 5.0], gamma=1.4),
                       cases.isentropic_vortex(1.4))]:
    g = BoxSolver(**kw)
    u = cases.to_state(ic(g.node_coords()), kw["gamma"])
    g.set_state(u)
    g.advance_to(0.0, 1e30, max_steps=64)
    g.synchronize()
    t0 = time.perf_counter()
    t, n = g.advance_to(0.0, 1e30, max_steps=4000)
    g.synchronize()
    dt = time.perf_counter() - t0
    print(f"{label}: {n / dt:9.0f} steps/s ({1e6 * dt / n:6.1f} us/step), graphs {'off' if os.environ.get('WARPII_GPU_NO_GRAPH') == '1' else 'on'}")
    g.close()
