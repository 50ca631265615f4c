"""Diagnostic: device point fluxes vs oracle on adjacent-node state pairs of a fine vortex mesh."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dgsem_cases as cases
import oracle
from oracle import Oracle
from warpii_b200.capi import point_fluxes
g = 1.4
o = Oracle(2, 3, [256, 256], [0.0, -5.0], [10.0, 5.0], gamma=g)
u = o.project(cases.isentropic_vortex(g))
rng = np.random.default_rng(1)
els = rng.integers(0, o.n_elems, 4000)
qa = u[els][:, :, 5]; qb = u[els][:, :, 6]
ec, es, prim = point_fluxes(qa, qb, 0, g)
ec_o = np.array([oracle.ec_flux(2, a, b, g)[:, 0] for a, b in zip(qa, qb)])
es_o = np.array([oracle.es_flux(2, a, b, [1.0, 0.0], g) for a, b in zip(qa, qb)])
np.set_printoptions(linewidth=200, precision=3)
print("EC max rel err per comp", np.abs(ec - ec_o).max(axis=0) / np.abs(ec_o).max(axis=0))
print("ES max rel err per comp", np.abs(es - es_o).max(axis=0) / np.abs(es_o).max(axis=0))
p_o = np.array([oracle.pressure(a, g) for a in qa]); beta_o = qa[:, 0] / (2.0 * p_o)
print("p bitwise equal:", np.array_equal(prim[:, 8], p_o), " beta bitwise equal:", np.array_equal(prim[:, 4], beta_o))
print("log rho  ulp diff hist:", np.unique(np.round((prim[:, 5] - np.log(qa[:, 0])) / np.spacing(np.abs(np.log(qa[:, 0])))), return_counts=True))
print("log beta ulp diff hist:", np.unique(np.round((prim[:, 6] - np.log(beta_o)) / np.spacing(np.abs(np.log(beta_o)))), return_counts=True))
inv = 1.0 / qa[:, 0]
print("u0 rel err", np.abs(prim[:, 1] - qa[:, 1] * inv).max(), " ib rel err", (np.abs(prim[:, 11] - 1.0 / beta_o) / (1.0 / beta_o)).max())
c = np.sqrt(g * p_o / qa[:, 0]); s = np.sqrt((qa[:, 1] * inv) ** 2 + (qa[:, 2] * inv) ** 2 + (qa[:, 3] * inv) ** 2)
print("lam rel err", (np.abs(prim[:, 10] - (s + c)) / (s + c)).max())
