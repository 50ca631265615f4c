"""Small run of the round-2 kernels (pencil stage kernel in 2-D / 3-D, Np = 3..5, one and two species, sources, fused field
phases and the stand-alone field kernel, walls / outflow, partial patches, the streamed host step), meant to be executed
under compute-sanitizer on the GPU box:
    compute-sanitizer --tool memcheck  python scripts/sanitize_pencil.py
    compute-sanitizer --tool racecheck python scripts/sanitize_pencil.py"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from warpii_b200 import BC_OUTFLOW, BC_WALL, BoxSolver  # noqa: E402

GAMMA = 5.0 / 3.0


def state(g, nsp, fields):
    xyz = g.node_coords()                                   # [elem][node][dim]
    dim = xyz.shape[-1]
    x = [xyz[..., d] for d in range(dim)] + [0.0 * xyz[..., 0]] * (3 - dim)
    s = np.sin(2 * np.pi * x[0]) * np.cos(2 * np.pi * x[1]) + 0.3 * np.sin(2 * np.pi * (x[2] + x[0]))
    kink = np.where(x[0] > 0.55, 1.0, 0.0)                  # a jump: the subcell blend becomes active somewhere
    nc = 5 * nsp + (8 if fields else 0)
    u = np.zeros((xyz.shape[0], nc, xyz.shape[1]))
    for sp in range(nsp):
        rho = (25.0 if sp == 0 and nsp == 2 else 1.0) * (1 + 0.1 * s + 0.5 * kink)
        vel = [0.2 * (1 + 0.2 * s), -0.1 * (1 + 0.1 * s), 0.05 + 0.0 * s]
        pr = 1.0 + 0.05 * s + 0.4 * kink
        u[:, 5 * sp, :] = rho
        for d in range(3):
            u[:, 5 * sp + 1 + d, :] = rho * vel[d]
        u[:, 5 * sp + 4, :] = pr / (GAMMA - 1) + 0.5 * rho * sum(v * v for v in vel)
    for c in range(8 if fields else 0):
        u[:, 5 * nsp + c, :] = 0.05 * (1 + c % 3) * np.cos(2 * np.pi * (x[0] + 0.2 * c)) * np.cos(2 * np.pi * x[1])
    return u


CASES = [  # dim, p, nx, periodic, nsp, fields (sources + Maxwell)
    (2, 3, [9, 5], True, 1, False),        # partial patches (8x4 patch on a 9x5 mesh)
    (2, 2, [4, 6], True, 1, False),
    (2, 4, [5, 4], False, 1, False),
    (3, 3, [3, 2, 3], True, 1, False),
    (3, 2, [4, 3, 2], False, 1, False),
    (3, 4, [2, 2, 3], True, 1, False),
    (2, 3, [8, 4], True, 2, True),         # fused field phases
    (3, 3, [2, 4, 2], True, 2, True),      # stand-alone field kernel
    (3, 2, [3, 3, 2], False, 2, True),
]
for dim, p, nx, periodic, nsp, fields in CASES:
    per = [int(periodic)] * dim
    bc = None if periodic else np.array([[[BC_WALL, BC_OUTFLOW][f % 2] for f in range(2 * dim)]] * nsp)
    g = BoxSolver(dim, p, nx, [0.0] * dim, [1.0, 1.2, 0.9][:dim], periodic=per, gamma=GAMMA, n_species=nsp, fields_enabled=fields,
                  n_boundaries=None if periodic else 2 * dim, bc_kinds=bc)
    if fields:
        g.set_sources(True, 1.3, 0.8, [0.04, -1.0][:nsp])
        g.set_maxwell(True, light_speed=2.5, chi=0.8, gamma=1.2)
    u = state(g, nsp, fields)
    g.upload(0, u)
    g.rhs(1, 0)
    dt = g.recommend_dt(0)
    g.ssprk2_step(dt, 0.0)
    g.advance_to(0.0, 1e9, max_steps=3)
    g.shock_indicator(0)
    host = g.download(0)
    if periodic:
        flat = np.ascontiguousarray(host.reshape(-1))
        g.host_step(flat, flat, dt, 0.0, n_slabs=3)
        host = flat
    assert np.isfinite(host).all()
    g.close()
    print("ok", dim, p, nx, "periodic" if periodic else "bounded", nsp, "fields" if fields else "", flush=True)
