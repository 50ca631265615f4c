"""Diagnostic: where is the GPU-vs-oracle energy RHS error, and who is closer to an extended-precision evaluation?"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dgsem_cases as cases
import oracle
from oracle import Oracle
from warpii_b200 import BoxSolver
np.set_printoptions(linewidth=220, precision=6)
n = 128; g = 1.4
o = Oracle(2, 3, [n, n], [0.0, -5.0], [10.0, 5.0], gamma=g, threads=16)
G = BoxSolver(2, 3, [n, n], [0.0, -5.0], [10.0, 5.0], gamma=g)
u = o.project(cases.isentropic_vortex(g))
G.upload_global(0, u); G.rhs(1, 0); got = G.download_global(1)
want, _ = o.rhs(u)
err = np.abs(got - want)[:, 4, :]
e, j = np.unravel_index(np.argmax(err), err.shape)
print("max energy err", err.max(), "elem", e, (e % n, e // n), "node", j, "got", got[e, 4, j], "want", want[e, 4, j])
print("errors at that elem (energy):", err[e])
print("alpha gpu max", G.shock_indicator_global(0).max(), "alpha oracle max", o.alpha(u).max())
# extended-precision volume+face evaluation at that node
L = np.longdouble
D = oracle.diff_matrix(4).astype(L); x, w = oracle.gll(4); w = w.astype(L)
h = [L(10.0) / n, L(10.0) / n]
def prim(q):
    q = q.astype(L); rho = q[0]; v = q[1:4] / rho; p = L(g - 1) * (q[4] - L(0.5) * rho * (v @ v)); return rho, v, p
def lnavg(a, b):
    z = (b - a) / (b + a)
    if abs(z) < 1e-3:
        s = L(1); zz = z * z; t = L(1)
        for k in range(1, 12): t *= zz; s += t / (2 * k + 1)
        return (a + b) / 2 / s
    return (b - a) / (np.log(b) - np.log(a))
def ec(qa, qb, d):
    ra, va, pa = prim(qa); rb, vb, pb = prim(qb); ba, bb = ra / (2 * pa), rb / (2 * pb)
    rl, bl = lnavg(ra, rb), lnavg(ba, bb); ravg = (ra + rb) / 2; uavg = (va + vb) / 2
    phat = ravg / (ba + bb); hh = 1 / (2 * bl * L(g - 1)) - L(0.25) * (va @ va + vb @ vb) + phat / rl + uavg @ uavg
    F = np.zeros(5, dtype=L); F[0] = rl * uavg[d]; F[1:4] = rl * uavg[d] * uavg; F[1 + d] += phat; F[4] = rl * uavg[d] * hh
    return F
def phys(q, d):
    r, v, p = prim(q); F = np.zeros(5, dtype=L); F[0] = q[1 + d]; F[1:4] = q[1:4].astype(L) * v[d]; F[1 + d] += p; F[4] = v[d] * (q[4] + p); return F
ue = u[e]; idx = (j % 4, j // 4)
r = np.zeros(5, dtype=L)
for d in range(2):
    st = 1 if d == 0 else 4; jd = idx[d]; base = j - jd * st
    acc = np.zeros(5, dtype=L)
    for l in range(4):
        F = phys(ue[:, j], d) if l == jd else ec(ue[:, j], ue[:, base + l * st], d)
        acc += D[jd, l] * F
    r += -2 / h[d] * acc
    # faces: traces are identical across faces for this smooth IC => f(u).n - f* = f(u)-F#(u,u) = 0 in exact arithmetic
print("extended precision (volume only; face terms vanish analytically):", r.astype(np.float64))
print("gpu   :", got[e, :, j]); print("oracle:", want[e, :, j])
