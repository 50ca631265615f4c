"""Low-storage Runge-Kutta on the GPU (rk.h:10-77): the fused stage of warpii_gpu_lsrk_stage and the
host layer's LowStorageRungeKuttaIntegrator against the oracle's restatement (run on the B200 box, -m gpu)."""
import numpy as np
import pytest

import dgsem_cases as cases
import mesh_cases as mc
from oracle import GeneralOracle, Oracle
from warpii_b200 import BC_OUTFLOW, BC_WALL, BoxSolver, WarpiiGpuError
from warpii_b200 import capi

pytestmark = pytest.mark.gpu


def rel(a, b):
    """relative L2 per component; a component that is (nearly) zero is measured against 1e-3 of the largest one"""
    norms = np.array([np.linalg.norm(b[:, c]) for c in range(b.shape[1])])
    return np.array([np.linalg.norm(a[:, c] - b[:, c]) / max(norms[c], 1e-3 * norms.max()) for c in range(b.shape[1])])


@pytest.mark.parametrize("dim,p,nx", [(1, 2, [24]), (2, 3, [12, 12]), (3, 2, [5, 4, 4])])
def test_stage_matches_oracle(dim, p, nx):
    left, right = [0.0, -5.0, -5.0][:dim], [10.0, 5.0, 5.0][:dim]
    o = Oracle(dim, p, nx, left, right, gamma=1.4, threads=4)
    g = BoxSolver(dim, p, nx, left, right, gamma=1.4)
    u = o.project(cases.isentropic_vortex(1.4) if dim > 1 else cases.sine_wave())
    r_in = u * (1.0 + 0.01 * np.cos(np.arange(u.size).reshape(u.shape)))
    # general stage: three different vectors
    g.upload_global(0, u)
    g.upload_global(1, r_in)
    g.lsrk_stage(sol_out=0, r_out=2, sol_in=0, r_in=1, factor_solution=2e-3, factor_ai=3e-3)
    sol, r_out = u.copy(), np.zeros_like(u)
    o.lsrk_stage(sol, r_out, r_in, 2e-3, 3e-3)
    assert (rel(g.download_global(0), sol) <= 1e-13).all() and (rel(g.download_global(2), r_out) <= 1e-13).all()
    # first stage of a step: r_in is the solution itself, so the new solution goes to a third vector
    g.upload_global(0, u)
    g.lsrk_stage(sol_out=2, r_out=1, sol_in=0, r_in=0, factor_solution=2e-3, factor_ai=1e-3)
    sol, r_out = u.copy(), np.zeros_like(u)
    o.lsrk_stage(sol, r_out, u.copy(), 2e-3, 1e-3)
    assert (rel(g.download_global(2), sol) <= 1e-13).all() and (rel(g.download_global(1), r_out) <= 1e-13).all()
    assert np.array_equal(g.download_global(0), u)   # the input is untouched
    # last stage: factor_ai = 0 leaves r_out alone
    before = g.download_global(3)
    g.lsrk_stage(sol_out=2, r_out=3, sol_in=2, r_in=1, factor_solution=1e-3, factor_ai=0.0)
    assert np.array_equal(g.download_global(3), before)
    # what must be refused: writing into the vector whose traces are being read
    for args in [(1, 2, 0, 1), (0, 1, 0, 1), (0, 0, 0, 1), (2, 0, 0, 1)]:
        with pytest.raises(WarpiiGpuError):
            g.lsrk_stage(*args, 1e-3, 1e-3)
    g.close()


@pytest.mark.parametrize("scheme", [0, 1, 2, 3])
def test_host_integrator_steps(scheme):
    b, a, c = capi.lsrk_coefficients(scheme)
    bc = [[BC_WALL, BC_OUTFLOW, BC_WALL, BC_WALL]]
    o = Oracle(2, 3, [10, 8], [0.0, -5.0], [10.0, 5.0], periodic=[0, 0], gamma=1.4, bc_kinds=bc, threads=4)
    g = BoxSolver(2, 3, [10, 8], [0.0, -5.0], [10.0, 5.0], periodic=[0, 0], gamma=1.4, n_boundaries=4, bc_kinds=bc)
    u = o.project(cases.isentropic_vortex(1.4))
    g.set_state_global(u)
    dt = 0.5 * o.recommend_dt(u)
    for k in range(10):
        g.lsrk_step(scheme, dt, k * dt)
        o.lsrk_step(u, b, a, c, dt, k * dt)
    assert (rel(g.get_state_global(), u) <= 1e-11).all()
    g.close()


def test_stage_on_curved_mesh():
    dim, p, nx, left, right = 2, 3, [8, 8], [0.0, 0.0], [1.0, 1.0]
    mesh, xyz = mc.mapped_box(dim, p, nx, left, right, [1, 1], mc.wavy(left, right, 0.05))
    geo = capi.mapped_metrics(dim, p, xyz, mesh["face_neighbor"])
    o = GeneralOracle(dim, p, mesh, geo, gamma=1.4, threads=4)
    g = capi.MeshSolver(dim, p, mesh, geo, gamma=1.4, n_vectors=3)
    prim = mc.periodic_state(1.4, left, right, dim)(mc.box_node_coords(dim, p, nx, left, right))
    mc.add_kinks(prim)
    u = mc.state_from(prim, 1.4)
    g.upload(0, u)
    g.lsrk_stage(sol_out=2, r_out=1, sol_in=0, r_in=0, factor_solution=1e-3, factor_ai=2e-3)
    sol, r_out = u.copy(), np.zeros_like(u)
    o.lsrk_stage(sol, r_out, u.copy(), 1e-3, 2e-3)
    assert (rel(g.download(2), sol) <= 1e-13).all() and (rel(g.download(1), r_out) <= 1e-13).all()
    g.close()
