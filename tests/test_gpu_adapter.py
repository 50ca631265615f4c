"""The reference's own SSPRK2Integrator (src/rk.h:79-117, compiled unmodified) driving the GPU operator through the adapter
header include/gpu_es_dgsem_operator.h on a deal.II stand-in (tests/adapter/adapter_check.cc): recommend_dt ->
evolve_one_time_step -> perform_forward_euler_step x 2, state in FiveMSolutionVec, DoF translation through
component_to_system_index / get_dof_indices.  The result must equal the same ABI calls made directly, bit for bit, and the
oracle to the parity tolerance."""
import numpy as np
import pytest

import adapter_check
import dgsem_cases as cases
import oracle
from oracle import Oracle
from warpii_b200 import BC_INFLOW, BC_OUTFLOW, BC_WALL, BoxSolver

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not adapter_check.available(), reason="adapter check library not built")]


def to_stub(u):      # [cell][comp][node] -> the stub DoFHandler's [cell][node][comp]
    return np.ascontiguousarray(np.transpose(u, (0, 2, 1)))


def from_stub(s):
    return np.ascontiguousarray(np.transpose(s, (0, 2, 1)))


@pytest.mark.parametrize("dim,p,nx,left,right,ic", [
    (2, 3, [9, 7], [0.0, -5.0], [10.0, 5.0], cases.isentropic_vortex(1.4)),
    (3, 2, [4, 3, 5], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], cases.smooth_blob_3d(0.1)),
    (1, 4, [12], [0.0], [1.0], cases.sine_wave()),
])
def test_reference_integrator_over_the_adapter_periodic(dim, p, nx, left, right, ic):
    gamma, steps = 1.4, 6
    o = Oracle(dim, p, nx, left, right, gamma=gamma, threads=4)
    u0 = o.project(ic)
    rc, err, st, t, _ = adapter_check.run(dim, p, nx, left, right, [1] * dim, to_stub(u0), steps, gamma)
    assert rc == 0, err
    got = from_stub(st)
    # the same calls made directly
    g = BoxSolver(dim, p, nx, left, right, gamma=gamma)
    g.upload_global(0, u0)
    tt = 0.0
    for _ in range(steps):
        dt = g.recommend_dt(0)
        g.ssprk2_step(dt, tt)
        tt += dt
    assert t == tt
    assert np.array_equal(got, g.download_global(0))
    g.close()
    u = u0.copy()
    assert o.solve(u, t, max_steps=steps) == steps
    err = cases.rel_l2_per_component(got, u)
    live = [c for c in range(5) if np.linalg.norm(u[:, c]) > 0]
    assert (err[live] <= 1e-11).all(), err


def test_reference_integrator_over_the_adapter_with_boundaries():
    """walls, supersonic outflow and an inflow Function<dim> (EulerBCMap from the reference's bc_helper.h), boundary-integrated
    fluxes in FiveMSolutionVec::boundary_integrated_fluxes"""
    dim, p, nx, left, right, gamma, steps = 2, 3, [6, 5], [0.0, 0.0], [1.0, 1.0], 1.4, 5
    bc = [[BC_INFLOW, BC_OUTFLOW, BC_WALL, BC_WALL]]
    q_in = oracle.primitive_to_conserved([1.1, 0.8, 0.05, 0.0, 1.2], gamma)
    inflow = np.zeros((1, 4, 5))
    inflow[0, 0] = q_in
    o = Oracle(dim, p, nx, left, right, periodic=[0, 0], gamma=gamma, bc_kinds=bc, threads=4)
    o.set_inflow(0, 0, q_in)
    u0 = o.project(cases.smooth_blob_3d(0.1))
    rc, err, st, t, bif = adapter_check.run(dim, p, nx, left, right, [0, 0], to_stub(u0), steps, gamma, bc_kind=bc, inflow=inflow)
    assert rc == 0, err
    u = u0.copy()
    bif_o = np.zeros(20)
    assert o.solve(u, t, bif=bif_o, max_steps=steps) == steps
    assert (cases.rel_l2_per_component(from_stub(st), u) <= 1e-11).all()
    assert np.allclose(bif, bif_o, rtol=1e-10, atol=1e-13)


def test_reference_integrator_over_the_adapter_with_subsonic_outflow():
    """EulerBCMap::set_subsonic_outflow_boundary (bc_helper.h:12-35) through the adapter: the ghost state keeps the interior
    density and momentum and takes the prescribed total energy (fluid_flux_es_dgsem_operator.h:385-390)."""
    from warpii_b200 import BC_SUBSONIC_OUTFLOW
    dim, p, nx, left, right, gamma, steps = 2, 2, [5, 4], [0.0, 0.0], [1.0, 1.0], 1.4, 5
    bc = [[BC_INFLOW, BC_SUBSONIC_OUTFLOW, BC_WALL, BC_WALL]]
    q_in = oracle.primitive_to_conserved([1.0, 0.3, 0.0, 0.0, 1.0], gamma)
    q_out = np.array([0.0, 0.0, 0.0, 0.0, 0.95 / (gamma - 1.0) + 0.5 * 0.09])     # only the energy is read
    inflow = np.zeros((1, 4, 5))
    inflow[0, 0], inflow[0, 1] = q_in, q_out
    o = Oracle(dim, p, nx, left, right, periodic=[0, 0], gamma=gamma, bc_kinds=bc, threads=4)
    o.set_inflow(0, 0, q_in)
    o.set_inflow(0, 1, q_out)
    u0 = o.project(cases.smooth_blob_3d(0.05))
    rc, err, st, t, bif = adapter_check.run(dim, p, nx, left, right, [0, 0], to_stub(u0), steps, gamma, bc_kind=bc, inflow=inflow)
    assert rc == 0, err
    u = u0.copy()
    bif_o = np.zeros(20)
    assert o.solve(u, t, bif=bif_o, max_steps=steps) == steps
    assert (cases.rel_l2_per_component(from_stub(st), u) <= 1e-11).all()
    assert np.allclose(bif, bif_o, rtol=1e-10, atol=1e-13)
