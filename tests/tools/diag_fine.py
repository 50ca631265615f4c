"""Diagnostic: GPU-vs-oracle RHS error on finer meshes (ln_avg conditioning grows as neighbouring states get closer)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dgsem_cases as cases
import oracle
from oracle import Oracle
from warpii_b200 import BoxSolver
np.set_printoptions(linewidth=200, precision=3)
for n in (16, 64, 128, 256):
    o = Oracle(2, 3, [n, n], [0.0, -5.0], [10.0, 5.0], gamma=1.4, threads=16)
    g = BoxSolver(2, 3, [n, n], [0.0, -5.0], [10.0, 5.0], gamma=1.4)
    u = o.project(cases.isentropic_vortex(1.4))
    g.upload_global(0, u); g.rhs(1, 0); got = g.download_global(1)
    want, _ = o.rhs(u)
    print(n, "rel L2", cases.rel_l2_per_component(got, want), "maxabs", np.abs(got - want).max(axis=(0, 2)), flush=True)
    g.close()
