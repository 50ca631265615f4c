"""GPU parity tests: the CUDA path through the C ABI against the CPU oracle (run on the B200 box, -m gpu).

Tolerances are the ones BASELINE.json's north_star states: one RHS evaluation relative L2 <= 1e-12 per
component, <= 1e-10 after 100 steps, conservation to round-off.  Integer-like outputs (nothing here) would
be bit-exact; everything on this path is FP64.
"""
import numpy as np
import pytest

import oracle
from oracle import Oracle
import dgsem_cases as cases
from warpii_b200 import BC_INFLOW, BC_OUTFLOW, BC_WALL, BoxSolver

pytestmark = pytest.mark.gpu

RHS_TOL = 1e-12       # north_star: one RHS evaluation, relative L2 per component
STEPS_TOL = 1e-10     # north_star: after 100 steps
DT_TOL = 1e-13        # recommend_dt (parity unpinned in the reference; oracle restatement)


def make_pair(dim, p, nx, left, right, periodic=None, gamma=1.4, n_species=1, fields=False, bc=None, threads=4):
    o = Oracle(dim, p, nx, left, right, periodic=periodic, gamma=gamma, n_species=n_species, fields_enabled=fields,
               bc_kinds=bc, threads=threads)
    nb = None
    if periodic is not None and not all(periodic):
        nb = 2 * dim
    g = BoxSolver(dim, p, nx, left, right, periodic=periodic, gamma=gamma, n_species=n_species, fields_enabled=fields,
                  n_boundaries=nb, bc_kinds=bc)
    assert g.shape == o.shape
    # single rank: the device numbers the elements patch by patch; l2g maps back to the oracle's lexicographic order
    assert sorted(g.l2g.tolist()) == list(range(o.n_elems))
    assert np.allclose(g.node_coords(), o.node_coords()[g.l2g], rtol=0, atol=1e-15)
    return o, g


def check_rhs(o, g, u, tol=RHS_TOL):
    g.upload_global(0, u)
    g.rhs(1, 0)
    got = g.download_global(1)
    want, _ = o.rhs(u)
    assert np.isfinite(got).all()
    # ||got - want|| <= max(1e-12 ||want||, two ulps of the terms that are differenced to form the RHS), per component: the
    # plain north_star tolerance wherever FP64 can deliver it, the rounding floor of the formula beyond (dgsem_cases.py)
    h = [(o_r - o_l) / n for o_l, o_r, n in zip(o.left, o.right, o.nx)]
    scale = cases.summand_scale(u, o.gamma, o.dim, h, oracle.diff_matrix(o.p + 1))
    err, bound = cases.rhs_error_and_bound(got, want, scale, tol)
    assert (err <= bound).all(), f"absolute L2 error per component {err}, bound {bound} (plain relative: {cases.rel_l2_per_component(got, want)})"
    return err


def check_rhs_plain(o, g, u, tol=RHS_TOL, zero_components=()):
    """The north_star criterion with nothing added: plain relative L2 <= 1e-12 on every component (those listed in
    zero_components have an exactly vanishing RHS in exact arithmetic and are checked against the differenced terms)."""
    g.upload_global(0, u)
    g.rhs(1, 0)
    got = g.download_global(1)
    want, _ = o.rhs(u)
    plain = cases.rel_l2_per_component(got, want)
    for c in range(want.shape[1]):
        if c in zero_components:
            continue
        assert plain[c] <= tol, f"component {c}: plain relative L2 {plain[c]} (all: {plain})"
    return plain


RHS_CASES = [
    # dim, p, nx, left, right, ic, gamma
    (1, 2, [20], [0.0], [1.0], cases.sine_wave(), 5.0 / 3.0),
    (1, 4, [7], [0.0], [1.0], cases.sine_wave(vel=(0.7, 0.2, -0.1)), 5.0 / 3.0),
    (1, 1, [16], [0.0], [1.0], cases.sine_wave(amp=0.3), 1.4),
    (2, 3, [16, 16], [0.0, -5.0], [10.0, 5.0], cases.isentropic_vortex(), 1.4),
    (2, 2, [9, 5], [0.0, 0.0], [1.0, 0.5], cases.sine_wave(vel=(1.0, 1.0, 0.0), wave=(1, 2, 0)), 5.0 / 3.0),
    (2, 5, [4, 6], [0.0, 0.0], [1.0, 1.0], cases.smooth_blob_3d(), 5.0 / 3.0),
    (3, 3, [5, 4, 6], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], cases.smooth_blob_3d(), 5.0 / 3.0),
    (3, 4, [4, 4, 4], [0.0, -5.0, -5.0], [10.0, 5.0, 5.0], cases.isentropic_vortex(), 1.4),
    (3, 2, [6, 6, 3], [0.0, 0.0, 0.0], [1.0, 2.0, 1.0], cases.smooth_blob_3d(0.1), 1.4),
    (3, 6, [2, 2, 2], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], cases.smooth_blob_3d(0.1), 5.0 / 3.0),
    # fine meshes: neighbouring states nearly equal => ln_avg and the flux differences are at their worst conditioning
    # (h of the 512^2 / 128^3 BASELINE configs is reached at 128^2 / 32^3 on a quarter / a 64th of the domain)
    (2, 3, [128, 128], [0.0, -5.0], [10.0, 5.0], cases.isentropic_vortex(), 1.4),
    (2, 3, [128, 128], [2.5, -1.25], [5.0, 1.25], cases.isentropic_vortex(), 1.4),
    (3, 4, [8, 8, 8], [4.0, -0.5, -0.5], [5.0, 0.5, 0.5], cases.isentropic_vortex(), 1.4),
    (3, 3, [16, 16, 16], [0.0, -5.0, -5.0], [10.0, 5.0, 5.0], cases.isentropic_vortex(), 1.4),
]


@pytest.mark.parametrize("dim,p,nx,left,right,ic,gamma", RHS_CASES)
def test_one_rhs_periodic(dim, p, nx, left, right, ic, gamma):
    o, g = make_pair(dim, p, nx, left, right, gamma=gamma)
    u = o.project(ic)
    check_rhs(o, g, u)
    # all of these meshes are coarse enough for the plain criterion (kappa <= 6e3); components that vanish identically
    # (momentum along a direction the flow does not depend on) are named
    vanishing = (3,) if ic.__qualname__.startswith("isentropic_vortex") else ()   # z-momentum of a flow extruded in z
    check_rhs_plain(o, g, u, zero_components=vanishing)
    # blending factors agree (all zero for these smooth states, but compare the tables anyway)
    assert np.allclose(g.shock_indicator_global(0), o.alpha(u), rtol=0, atol=1e-12)
    g.close()


def test_sod_rhs_with_shock_capturing_and_outflow():
    """BASELINE config 1: 1D Sod, p=2, 200 cells, Outflow at both ends; alpha > 0 at the discontinuity."""
    gamma = 5.0 / 3.0
    bc = [[BC_OUTFLOW, BC_OUTFLOW]]
    o, g = make_pair(1, 2, [200], [0.0], [1.0], periodic=[0], gamma=gamma, bc=bc)
    u = o.project(cases.sod())
    a_ref = o.alpha(u)
    assert a_ref.max() > 0
    g.upload_global(0, u)
    a_gpu = g.shock_indicator_global(0)
    assert np.allclose(a_gpu, a_ref, rtol=1e-10, atol=1e-12)
    check_rhs(o, g, u)
    # after some steps the profile has a rarefaction, contact and shock: compare again there
    o.solve(u, 0.02)
    check_rhs(o, g, u)
    g.close()


def test_sod_full_run_config1():
    gamma = 5.0 / 3.0
    bc = [[BC_OUTFLOW, BC_OUTFLOW]]
    o, g = make_pair(1, 2, [200], [0.0], [1.0], periodic=[0], gamma=gamma, bc=bc)
    u = o.project(cases.sod())
    g.set_state_global(u)
    steps_g = g.solve(0.1)
    steps_o = o.solve(u, 0.1)
    got = g.get_state_global()
    assert steps_g == steps_o
    # A shock run is chaotic in the last bits (alpha switches); the profiles must still agree closely.
    err = cases.rel_l2_per_component(got, u)
    assert err[0] < 1e-8 and err[1] < 1e-8 and err[4] < 1e-8, err
    g.close()


def test_inflow_outflow_bif_and_balance():
    """test/conservation_test.cc:85-137 through the GPU path, and parity of the boundary-integrated fluxes."""
    gamma = 1.6666666666667
    bc = [[BC_INFLOW, BC_OUTFLOW]]
    o, g = make_pair(1, 4, [1], [0.0], [1.0], periodic=[0], gamma=gamma, bc=bc)
    q_in = oracle.primitive_to_conserved([3.857, 2.629, 0.0, 0.0, 10.333], gamma)
    o.set_inflow(0, 0, q_in)
    g.set_inflow(0, 0, q_in)
    u = o.project(cases.sine_wave())
    ic = o.global_integral(u)
    g.set_state_global(u)
    assert np.allclose(g.global_integral(0), ic, rtol=0, atol=1e-15)
    steps = g.solve(0.04)
    assert steps > 0
    bif = g.boundary_fluxes(0)
    now = g.global_integral(0)
    balance = now + bif[0:5] + bif[5:10]
    for c in range(3):
        assert abs(balance[c] - ic[c]) < 1e-14
    bif_o = np.zeros(10)
    steps_o = o.solve(u, 0.04, bif=bif_o)
    assert steps_o == steps
    assert np.allclose(bif, bif_o, rtol=1e-11, atol=1e-13)
    assert (cases.rel_l2_per_component(g.get_state_global(), u)[[0, 1, 4]] < STEPS_TOL).all()
    g.close()


def test_walls_2d_kelvin_helmholtz_rhs():
    """Walls in x, periodic in y (the reference KH set-up): Gauss(p+2) boundary-face path in 2D."""
    gamma = 5.0 / 3.0
    k = 1.2 * np.pi
    bc = [[BC_WALL, BC_WALL, BC_WALL, BC_WALL]]
    o, g = make_pair(2, 3, [12, 10], [-0.5, 0.0], [0.5, 2 * np.pi / k], periodic=[0, 1], gamma=gamma, bc=bc)
    u = o.project(cases.kelvin_helmholtz(k))
    check_rhs(o, g, u)
    a_ref = o.alpha(u)
    assert np.allclose(g.shock_indicator_global(0), a_ref, rtol=1e-9, atol=1e-12)
    g.close()


@pytest.mark.parametrize("dim,p,nx", [(1, 3, [12]), (2, 2, [6, 5])])
def test_subsonic_outflow_ghost_state(dim, p, nx):
    """EulerBCMap::set_subsonic_outflow_boundary (bc_helper.h:12-35): ghost = inside state with the total energy replaced by
    the prescribed one (fluid_flux_es_dgsem_operator.h:385-390).  No input file of the reference can select it
    (species.cc:48-49 maps "Outflow" to the supersonic kind), the operator and the ABI know it."""
    from warpii_b200 import BC_SUBSONIC_OUTFLOW
    gamma = 1.4
    kinds = [BC_INFLOW, BC_SUBSONIC_OUTFLOW] + [BC_WALL] * (2 * dim - 2)
    o, g = make_pair(dim, p, nx, [0.0] * dim, [1.0] * dim, periodic=[0] * dim, gamma=gamma, bc=[kinds])
    q_in = oracle.primitive_to_conserved([1.0, 0.3, 0.0, 0.0, 1.0], gamma)
    q_out = np.array([0.0, 0.0, 0.0, 0.0, 2.4])     # only the energy is read
    for s in (o, g):
        s.set_inflow(0, 0, q_in)
        s.set_inflow(0, 1, q_out)
    u = o.project(cases.sine_wave(amp=0.1, vel=(0.3, 0.0, 0.0)))
    check_rhs(o, g, u)
    want, bif_o = o.rhs(u)
    # the outflow face really used the prescribed energy: with the supersonic kind the answer differs
    o2, g2 = make_pair(dim, p, nx, [0.0] * dim, [1.0] * dim, periodic=[0] * dim, gamma=gamma, bc=[[BC_INFLOW, BC_OUTFLOW] + [BC_WALL] * (2 * dim - 2)])
    o2.set_inflow(0, 0, q_in)
    other, _ = o2.rhs(u)
    assert np.abs(other - want).max() > 1e-3
    g.close()
    g2.close()


def test_mixed_bcs_3d_rhs():
    gamma = 1.4
    bc = [[BC_INFLOW, BC_OUTFLOW, BC_WALL, BC_WALL, BC_OUTFLOW, BC_WALL]]
    o, g = make_pair(3, 2, [4, 3, 3], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], periodic=[0, 0, 0], gamma=gamma, bc=bc)
    q_in = oracle.primitive_to_conserved([1.1, 0.4, 0.1, -0.05, 1.2], gamma)
    o.set_inflow(0, 0, q_in)
    g.set_inflow(0, 0, q_in)
    u = o.project(cases.smooth_blob_3d(0.1))
    check_rhs(o, g, u)
    g.close()


def test_two_species_with_fields_rhs_and_step():
    """nc = 18: two fluids + 8 field components that the operator carries through unchanged."""
    gamma = 5.0 / 3.0
    o, g = make_pair(2, 3, [6, 5], [0.0, 0.0], [1.0, 1.0], gamma=gamma, n_species=2, fields=True)
    u = o.project(cases.sine_wave(vel=(0.5, 0.3, 0.1), wave=(1, 1, 0)), species=0)
    u = o.project(cases.smooth_blob_3d(0.15), species=1, u=u)
    rng = np.random.default_rng(12345)
    u[:, 10:18, :] = rng.standard_normal(u[:, 10:18, :].shape)
    check_rhs(o, g, u)
    dt = 0.3 * o.recommend_dt(u)
    g.upload_global(0, u)
    g.ssprk2_step(dt, 0.0)
    got = g.download_global(0)
    want = u.copy()
    o.ssprk2_step(want, dt, 0.0)
    assert (cases.rel_l2_per_component(got, want) < 1e-13).all()
    assert np.array_equal(got[:, 10:18, :], u[:, 10:18, :])   # fields: 0.5*u + 0.5*(u + dt*0) exactly
    g.close()


@pytest.mark.parametrize("dim,p,nx,left,right,ic,gamma", [RHS_CASES[0], RHS_CASES[3], RHS_CASES[6]])
def test_recommend_dt_parity(dim, p, nx, left, right, ic, gamma):
    o, g = make_pair(dim, p, nx, left, right, gamma=gamma)
    u = o.project(ic)
    g.upload_global(0, u)
    want = o.recommend_dt(u)
    got = g.recommend_dt(0)
    assert abs(got - want) <= DT_TOL * want
    # fused CFL of the second SSPRK2 stage == stand-alone sweep of the same vector
    dt = 0.5 * want
    g.ssprk2_step(dt, 0.0)
    fused = g.recommend_dt(0)
    unew = g.download(0)
    g.upload(0, unew)          # invalidates the cached reduction
    assert g.recommend_dt(0) == fused
    o.ssprk2_step(u, dt, 0.0)
    assert abs(fused - o.recommend_dt(u)) <= 1e-12 * fused
    g.close()


def test_hundred_steps_vortex_2d():
    """north_star: <= 1e-10 after 100 steps, mass/momentum/energy conserved to round-off."""
    gamma = 1.4
    o, g = make_pair(2, 3, [16, 16], [0.0, -5.0], [10.0, 5.0], gamma=gamma, threads=8)
    u = o.project(cases.isentropic_vortex(gamma))
    ic = o.global_integral(u)
    g.set_state_global(u)
    t, steps = g.advance_to(0.0, 1e9, max_steps=100)
    assert steps == 100
    so = o.solve(u, t, max_steps=100)
    assert so == 100
    got = g.get_state_global()
    err = cases.rel_l2_per_component(got, u)
    assert (err[[0, 1, 2, 4]] <= STEPS_TOL).all(), err
    now = g.global_integral(0)
    for c in (0, 1, 2, 4):
        assert abs(now[c] - ic[c]) <= 1e-12 * max(1.0, abs(ic[c])), (c, now[c], ic[c])
    g.close()


def test_hundred_steps_3d():
    gamma = 5.0 / 3.0
    o, g = make_pair(3, 3, [4, 4, 4], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], gamma=gamma, threads=8)
    u = o.project(cases.smooth_blob_3d(0.1))
    ic = o.global_integral(u)
    g.set_state_global(u)
    t, steps = g.advance_to(0.0, 1e9, max_steps=100)
    o.solve(u, t, max_steps=100)
    err = cases.rel_l2_per_component(g.get_state_global(), u)
    assert (err <= STEPS_TOL).all(), err
    now = g.global_integral(0)
    for c in range(5):
        assert abs(now[c] - ic[c]) <= 1e-12 * max(1.0, abs(ic[c]))
    g.close()


def test_reference_periodic_1d_conservation():
    """test/conservation_test.cc:45-83 through the GPU path: one p=4 cell, periodic onto itself."""
    o, g = make_pair(1, 4, [1], [0.0], [1.0], gamma=1.6666666666667)
    u = o.project(cases.sine_wave())
    g.set_state_global(u)
    ic = g.global_integral(0)
    assert g.solve(0.04) > 0
    now = g.global_integral(0)
    for c in range(3):
        assert abs(now[c] - ic[c]) < 1e-13
    g.close()


def test_reference_freestream_1d_convergence():
    """test/input_test.cc:18-66 through the GPU path."""
    from test_oracle_golden import _l2_error_density
    errs = []
    for nx in (20, 30):
        o, g = make_pair(1, 2, [nx], [0.0], [1.0], gamma=1.6666666666667)
        g.set_state_global(o.project(cases.sine_wave()))
        g.solve(0.04)
        exact = lambda x: 1 + 0.6 * np.sin(2 * np.pi * (x - 0.04))
        errs.append(_l2_error_density(o, g.get_state_global(), exact, 2))
        # the product's own compute_global_error (dg_solution_helper.cc:50-69) gives the same number
        mine = g.global_error(lambda x: [exact(x[0]), 0, 0, 0, 0], component=0)
        assert abs(mine - errs[-1]) <= 1e-11 * errs[-1]   # (two summation orders of the same quadrature)
        g.close()
    assert abs(errs[1]) < 1e-4
    assert abs(errs[0] / errs[1] - 1.5 ** 3) < 1.0


def test_global_error_in_2d_against_numpy():
    o, g = make_pair(2, 3, [6, 5], [0.0, -1.0], [2.0, 1.0], gamma=1.4)
    u = o.project(cases.isentropic_vortex(1.4))
    g.set_state_global(u)
    exact = lambda x: np.array([1.0 + 0.1 * x[0], 0.2, 0.3 * x[1], 0.0, 2.5 + 0.05 * x[0] * x[1]])
    xq, wq = oracle.gauss(3)
    x, _ = oracle.gll(4)
    I = np.array([[np.prod([(xq[q] - x[m]) / (x[i] - x[m]) for m in range(4) if m != i]) for i in range(4)] for q in range(3)])
    h = [2.0 / 6, 2.0 / 5]
    for comp in (0, 2, 4):
        err2 = 0.0
        for e in range(o.n_elems):
            ex, ey = e % 6, e // 6
            ue = u[e, comp].reshape(4, 4)           # [y][x]
            uq = I @ ue @ I.T                       # [qy][qx]
            for qy in range(3):
                for qx in range(3):
                    pt = np.array([0.0 + (ex + xq[qx]) * h[0], -1.0 + (ey + xq[qy]) * h[1]])
                    err2 += (uq[qy, qx] - exact(pt)[comp]) ** 2 * wq[qx] * wq[qy] * h[0] * h[1]
        assert abs(g.global_error(exact, component=comp) - np.sqrt(err2)) <= 1e-12 * np.sqrt(err2)
    g.close()


def test_solver_callbacks_and_step_count():
    """FiveMomentDGSolver::solve + advance(): the writeout callback fires at the frame times and at t_end."""
    o, g = make_pair(1, 2, [10], [0.0], [1.0], gamma=1.6666666666667)
    g.set_state_global(o.project(cases.sine_wave()))
    times = []
    steps = g.solve(0.05, callback=times.append, callback_interval=0.01)
    assert steps >= 5
    assert len(times) == 5 and abs(times[-1] - 0.05) < 1e-12 and abs(times[0] - 0.01) < 1e-12
    g.close()


def test_error_conventions():
    from warpii_b200 import WarpiiGpuError
    o, g = make_pair(1, 2, [4], [0.0], [1.0])
    with pytest.raises(WarpiiGpuError):
        g.forward_euler_step(0, 0, 0.1, 0.0)          # dst must differ from u
    with pytest.raises(WarpiiGpuError):
        g.forward_euler_step(0, 7, 0.1, 0.0)          # unknown vector
    g.close()
    with pytest.raises(WarpiiGpuError):                # non-periodic box without boundary conditions declared
        BoxSolver(1, 2, [4], [0.0], [1.0], periodic=[0], n_boundaries=0)


def test_unphysical_state_is_reported():
    """The GPU analogue of --enable-fpe (fpe.cc:19-22): a negative pressure yields a NaN transport speed, and the time
    loop stops with an error instead of marching on."""
    from warpii_b200 import WarpiiGpuError
    o, g = make_pair(2, 3, [8, 8], [0.0, -5.0], [10.0, 5.0], gamma=1.4)
    u = o.project(cases.isentropic_vortex(1.4))
    bad = u.copy()
    bad[13, 4, 5] = 0.01          # total energy below the kinetic energy at one node
    g.upload_global(0, bad)
    assert np.isnan(g.max_transport_speed(0))
    with pytest.raises(WarpiiGpuError):
        g.advance_to(0.0, 1.0, max_steps=3)
    # the fused reduction of the second stage reports it too
    g.upload_global(0, u)
    g.ssprk2_step(0.9, 0.0)       # far beyond the stable step: the state blows up
    g.ssprk2_step(0.9, 0.9)
    v = g.max_transport_speed(0)
    assert not np.isfinite(v) or v > 1e3
    g.close()


# ---- full-size checks on the bench workload (size-independent properties; the oracle would take minutes here) ----
def test_full_size_c2_properties():
    """BASELINE config 2 (512x512, p=3): conservation over steps, free-stream preservation, translation invariance."""
    gamma = 1.4
    n = 512
    g = BoxSolver(2, 3, [n, n], [0.0, -5.0], [10.0, 5.0], gamma=gamma)
    xyz = g.node_coords()
    u0 = np.empty(g.shape)
    u0[g.l2g] = cases.to_state(cases.isentropic_vortex(gamma)(xyz), gamma)   # global (lexicographic) element order
    g.set_state_global(u0)
    ic = g.global_integral(0)
    t, steps = g.advance_to(0.0, 1e9, max_steps=10)
    assert steps == 10
    now = g.global_integral(0)
    for c in (0, 1, 2, 4):
        assert abs(now[c] - ic[c]) <= 2e-12 * max(1.0, abs(ic[c])), (c, now[c] - ic[c])
    u10 = g.get_state_global()
    assert np.isfinite(u10).all()
    # translation invariance: shifting the initial state by 64 elements in x and 32 in y shifts the answer, bit for bit
    sx, sy = 67, 33   # not multiples of the 4x4 patches: faces change between block-internal and block-external
    shifted = np.roll(np.roll(u0.reshape(n, n, 5, 16), sy, axis=0), sx, axis=1).reshape(u0.shape)
    g.set_state_global(shifted)
    g.advance_to(0.0, 1e9, max_steps=10)
    u10s = g.get_state_global()
    back = np.roll(np.roll(u10s.reshape(n, n, 5, 16), -sy, axis=0), -sx, axis=1).reshape(u0.shape)
    assert np.array_equal(back, u10)
    # free stream: a uniform state has zero residual up to round-off of the flux differences
    uni = np.zeros_like(u0)
    uni[:, :, :] = cases.primitive_to_conserved(np.array([1.3, 0.4, -0.7, 0.1, 0.9]), gamma)[None, :, None]
    g.upload(0, uni)
    g.rhs(1, 0)
    r = g.download(1)
    assert np.abs(r).max() < 1e-9   # |D| / h ~ 3e2, flux ~ 1, eps ~ 1e-16
    g.close()


def test_graph_replay_of_the_time_loop_is_bitwise_identical(monkeypatch):
    """Full batches of the device-resident loop are replayed from a CUDA graph; a context created with
    WARPII_GPU_NO_GRAPH=1 launches the same kernels one by one.  Same numbers, and a captured batch is dropped when a
    setter changes kernel arguments (here: the inflow state between two runs)."""
    gamma = 1.4
    bc = [[BC_INFLOW, BC_OUTFLOW]]
    q_a = oracle.primitive_to_conserved([1.0, 0.8, 0.0, 0.0, 1.0], gamma)
    q_b = oracle.primitive_to_conserved([1.3, 0.9, 0.0, 0.0, 1.2], gamma)
    results = []
    for no_graph in ("0", "1"):
        monkeypatch.setenv("WARPII_GPU_NO_GRAPH", no_graph)
        o, g = make_pair(1, 3, [24], [0.0], [1.0], periodic=[0], gamma=gamma, bc=bc)
        g.set_inflow(0, 0, q_a)
        u = o.project(cases.sine_wave(amp=0.2, vel=(0.8, 0.0, 0.0)))
        g.set_state_global(u)
        t, n1 = g.advance_to(0.0, 1e30, max_steps=40)          # 5 full batches
        g.set_inflow(0, 0, q_b)                                # must invalidate the captured batch
        t, n2 = g.advance_to(t, 1e30, max_steps=21)            # 2 full batches + 5 single steps
        results.append((t, n1 + n2, g.get_state_global(), g.boundary_fluxes(0), g.launch_count()))
        if no_graph == "1":
            o.set_inflow(0, 0, q_a)
            o.solve(u, 1e30, max_steps=40)
            o.set_inflow(0, 0, q_b)
            o.solve(u, 1e30, max_steps=21)
            assert (cases.rel_l2_per_component(results[-1][2], u)[[0, 1, 4]] < STEPS_TOL).all()
        g.close()
    (ta, na, ua, ba, la), (tb, nb, ub, bb, lb) = results
    assert ta == tb and na == nb == 61 and la == lb
    assert np.array_equal(ua, ub) and np.array_equal(ba, bb)


def blast(dim, width=0.02):
    """A pressure / density jump across a slanted plane: under-resolved on the test meshes, so the Persson-Peraire blend
    switches the subcell finite-volume fluxes on in the elements it crosses (in every direction of the mesh)."""
    def fn(xyz):
        s = xyz[..., 0] - 0.45 + (0.3 * (xyz[..., 1] - 0.5) if dim > 1 else 0.0) + (0.2 * (xyz[..., 2] - 0.5) if dim > 2 else 0.0)
        w = 0.5 * (1 - np.tanh(s / width))
        out = np.zeros(xyz.shape[:-1] + (5,))
        out[..., 0] = 0.125 + 0.875 * w
        out[..., 1] = 0.3 * w
        out[..., 2] = -0.2 * w
        out[..., 3] = 0.1 * w
        out[..., 4] = 0.1 + 0.9 * w
        return out
    return fn


@pytest.mark.parametrize("dim,p,nx", [(1, 1, [24]), (1, 5, [9]), (2, 1, [10, 9]), (2, 2, [8, 7]), (2, 3, [8, 8]), (2, 4, [5, 6]),
                                      (2, 5, [4, 4]), (3, 1, [6, 5, 4]), (3, 2, [5, 4, 4]), (3, 3, [4, 4, 3]), (3, 4, [3, 3, 3])])
def test_rhs_with_active_subcell_fv_blend(dim, p, nx):
    """Volume + FV blend in every dimension and for even / odd Np (full and half pair classes, adjacent-pair slots)."""
    o, g = make_pair(dim, p, nx, [0.0] * dim, [1.0] * dim, gamma=1.4)
    u = o.project(blast(dim))
    alpha = o.alpha(u)
    assert (alpha > 0).sum() >= 2 and (alpha == 0).sum() >= 1, alpha.ravel()
    check_rhs(o, g, u)
    assert np.allclose(g.shock_indicator_global(0), alpha, rtol=1e-9, atol=1e-12)
    # and a few steps through the blend (alpha changes from step to step)
    g.set_state_global(u)
    dt = 0.2 * g.recommend_dt(0)
    steps = g.solve(5 * dt, fixed_dt=dt)
    assert steps == 5 == o.solve(u, 5 * dt, fixed_dt=dt)
    err = cases.rel_l2_per_component(g.get_state_global(), u)
    assert (err < 1e-9).all(), err       # the logistic blend amplifies last-bit differences of the modal energies
    g.close()


@pytest.mark.parametrize("dim,p,nx,slabs", [(2, 3, [32, 24], 0), (2, 3, [32, 24], 5), (3, 2, [6, 5, 7], 3), (1, 4, [40], 7), (2, 2, [4, 4], 16)])
def test_streamed_host_step_is_the_plain_step_bit_for_bit(dim, p, nx, slabs):
    """warpii_gpu_host_ssprk2_step (upload, stages and download overlapped slab by slab) == upload + ssprk2_step + download."""
    left, right = [0.0, -5.0, -5.0][:dim], [10.0, 5.0, 5.0][:dim]
    o, g = make_pair(dim, p, nx, left, right, gamma=1.4)
    u = o.project(cases.isentropic_vortex(1.4) if dim > 1 else cases.sine_wave())[g.l2g].copy()
    g.upload(0, u)
    dt = g.recommend_dt(0)
    want, dts = u.copy(), []
    for k in range(3):
        g.upload(0, want)
        g.ssprk2_step(dt if k == 0 else dts[-1], 0.0)
        want = g.download(0)
        dts.append(g.recommend_dt(0))
    host = u.copy()
    got_dts, d = [], dt
    for k in range(3):
        d = g.host_step(host, host, d, 0.0, n_slabs=slabs)
        got_dts.append(d)
    assert np.array_equal(host, want)
    assert got_dts == dts
    # pinned buffers: the step is captured into a CUDA graph on the first call and replayed afterwards
    import torch
    pinned = torch.empty(u.size, dtype=torch.float64).pin_memory().numpy().reshape(u.shape)
    pinned[...] = u
    got_dts, d = [], dt
    for k in range(3):
        d = g.host_step(pinned, pinned, d, 0.0, n_slabs=slabs)
        got_dts.append(d)
    assert np.array_equal(pinned, want)
    assert got_dts == dts
    g.close()


def test_streamed_host_step_falls_back_with_boundaries():
    bc = [[BC_WALL, BC_OUTFLOW, BC_WALL, BC_WALL]]
    o, g = make_pair(2, 3, [10, 8], [0.0, -5.0], [10.0, 5.0], periodic=[0, 0], gamma=1.4, bc=bc)
    u = o.project(cases.isentropic_vortex(1.4))[g.l2g].copy()
    g.upload(0, u)
    dt = g.recommend_dt(0)
    g.ssprk2_step(dt, 0.0)
    want, want_dt = g.download(0), g.recommend_dt(0)
    host = u.copy()
    out = np.zeros_like(host)
    got_dt = g.host_step(host, out, dt, 0.0)
    assert np.array_equal(out, want) and got_dt == want_dt and np.array_equal(host, u)
    g.close()
