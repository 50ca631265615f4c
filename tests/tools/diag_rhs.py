"""Diagnostic: print GPU-vs-oracle RHS errors per component for the parity cases."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dgsem_cases as cases
from oracle import Oracle
from warpii_b200 import BoxSolver
from test_gpu_parity import RHS_CASES
np.set_printoptions(linewidth=200, precision=3)
for (dim, p, nx, left, right, ic, gamma) in RHS_CASES:
    o = Oracle(dim, p, nx, left, right, gamma=gamma, threads=4)
    g = BoxSolver(dim, p, nx, left, right, gamma=gamma)
    u = o.project(ic)
    g.upload(0, u); g.rhs(1, 0); got = g.download(1)
    want, _ = o.rhs(u)
    err = cases.rel_l2_per_component(got, want)
    nrm = np.array([np.linalg.norm(want[:, c, :]) for c in range(5)])
    absd = np.array([np.abs(got[:, c, :] - want[:, c, :]).max() for c in range(5)])
    print(dim, p, nx, "rel", err, "norm", nrm, "maxabs", absd, flush=True)
    g.close()
