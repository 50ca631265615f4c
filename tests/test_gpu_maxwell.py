"""GPU parity of the field system (perfectly hyperbolic Maxwell fluxes + two-fluid sources) against the oracle's definition
(tests/test_maxwell_cpu.py pins that definition).  New physics: the reference evolves nothing in the field components, so
this is GPU-vs-oracle only.  Tolerances as for the fluid: 1e-12 per RHS component (dgsem_cases.rhs_error_and_bound), 1e-13
on recommend_dt, 1e-10 over 100 steps."""
import numpy as np
import pytest

import dgsem_cases as cases
import oracle
from oracle import Oracle
from warpii_b200 import BC_OUTFLOW, BC_WALL, BoxSolver

pytestmark = pytest.mark.gpu

SRC = dict(epsilon0=1.3, chi=0.8, charge_over_mass=[0.04, -1.0])
MX = dict(light_speed=2.5, chi=0.8, gamma=1.2)


def two_fluid_state(o, seed=7, field_amp=0.1):
    xyz = o.node_coords()
    dim = o.dim
    x = [xyz[..., d] for d in range(dim)] + [0.0, 0.0]
    s = np.sin(2 * np.pi * x[0]) * np.cos(2 * np.pi * x[1]) + 0.3 * np.sin(2 * np.pi * (x[2] + x[0]))
    u = np.zeros(o.shape)
    for sp, (rho0, vel, p0) in enumerate([(25.0, (0.05, -0.02, 0.01), 1.0), (1.0, (-0.2, 0.1, 0.05), 1.0)][:o.nsp]):
        prim = np.zeros(xyz.shape[:-1] + (5,))
        prim[..., 0] = rho0 * (1 + 0.1 * s)
        for d in range(3):
            prim[..., 1 + d] = vel[d] * (1 + 0.2 * s)
        prim[..., 4] = p0 * (1 + 0.05 * s)
        cases.to_state(prim, o.gamma, nc=o.nc, species=sp, u=u)
    k = 5 * o.nsp
    for c, amp in enumerate([0.3, -0.2, 0.15, 0.1, -0.05, 0.2, 0.02, -0.03]):
        u[:, k + c, :] = field_amp * amp * (1 + 0.5 * np.cos(2 * np.pi * (x[0] + 0.3 * c)) * np.cos(2 * np.pi * x[1] * (1 + c % 2)))
    return u


def make(dim, p, nx, nsp=2, periodic=None, bc=None, sources=True):
    left, right = [0.0] * dim, [1.0] * dim
    kw = dict(gamma=5.0 / 3.0, n_species=nsp, fields_enabled=True)
    o = Oracle(dim, p, nx, left, right, periodic=periodic, bc_kinds=bc, threads=8, **kw)
    nb = None if periodic is None or all(periodic) else 2 * dim
    g = BoxSolver(dim, p, nx, left, right, periodic=periodic, bc_kinds=bc, n_boundaries=nb, **kw)
    qm = SRC["charge_over_mass"][:nsp]
    if sources:
        o.set_sources(True, SRC["epsilon0"], SRC["chi"], qm)
        g.set_sources(True, SRC["epsilon0"], SRC["chi"], qm)
    o.set_maxwell(True, **MX)
    g.set_maxwell(True, **MX)
    return o, g


def check(o, g, u):
    g.upload_global(0, u)
    g.rhs(1, 0)
    got = g.download_global(1)
    want, _ = o.rhs(u)
    h = [(r - l) / n for l, r, n in zip(o.left, o.right, o.nx)]
    D = oracle.diff_matrix(o.p + 1)
    scale = cases.summand_scale(u, o.gamma, o.dim, h, D) + cases.field_summand_scale(u, o.nsp, o.dim, h, D, **MX)
    err, bound = cases.rhs_error_and_bound(got, want, scale)
    assert np.isfinite(got).all()
    assert (err <= bound).all(), f"abs L2 error {err}, bound {bound}, plain {cases.rel_l2_per_component(got, want)}"
    plain = cases.rel_l2_per_component(got, want)
    assert (plain[5 * o.nsp:] <= 1e-12).all(), plain      # the field components: plain criterion on these coarse meshes


@pytest.mark.parametrize("dim,p,nx,nsp", [(1, 2, [12], 2), (1, 4, [5], 1), (2, 3, [6, 5], 2), (2, 2, [5, 4], 2), (2, 5, [3, 3], 1),
                                          (3, 3, [4, 3, 2], 2), (3, 2, [3, 3, 3], 2), (3, 4, [2, 2, 3], 1),
                                          # the degrees the node-per-thread stage kernel serves (Np = 2, 6, 7), stand-alone field kernel
                                          (2, 1, [7, 6], 2), (3, 1, [4, 3, 3], 2), (2, 6, [3, 2], 1), (3, 5, [2, 2, 2], 2), (3, 6, [2, 1, 2], 1)])
def test_rhs_with_fields_evolving(dim, p, nx, nsp):
    o, g = make(dim, p, nx, nsp)
    u = two_fluid_state(o)
    check(o, g, u)
    # the transport speed now also covers c max(1, chi, gamma) and the plasma / cyclotron frequency
    g.upload_global(0, u)
    want = o.recommend_dt(u)
    assert abs(g.recommend_dt(0) - want) <= 1e-13 * want
    g.close()


def test_strong_field_limits_dt_through_the_cyclotron_frequency():
    o, g = make(2, 3, [4, 4])
    u = two_fluid_state(o, field_amp=400.0)
    g.upload_global(0, u)
    want = o.recommend_dt(u)
    vfield = 16 * 0.5 / want
    assert vfield > 4 * 4 * MX["light_speed"] * 1.2 * 1.5     # the frequency bound, not the wave speed, sets dt here
    assert abs(g.recommend_dt(0) - want) <= 1e-13 * want
    g.close()


def test_zero_gradient_fields_at_walls():
    bc = [[BC_WALL, BC_OUTFLOW, BC_WALL, BC_WALL]] * 2
    o, g = make(2, 3, [5, 4], periodic=[0, 0], bc=bc)
    u = two_fluid_state(o)
    check(o, g, u)
    g.close()


@pytest.mark.parametrize("dim,p,nx", [(2, 3, [6, 6]), (3, 3, [3, 3, 3]), (1, 3, [10])])
def test_hundred_steps_with_fields(dim, p, nx):
    o, g = make(dim, p, nx)
    u = two_fluid_state(o)
    g.set_state_global(u)
    t, steps = g.advance_to(0.0, 1e9, max_steps=100)
    assert steps == 100
    assert o.solve(u, t, max_steps=100) == 100
    err = cases.rel_l2_per_component(g.get_state_global(), u)
    assert (err <= 1e-10).all(), err
    # the fused reduction of the last stage == the stand-alone sweep == the oracle's
    fused = g.recommend_dt(0)
    assert abs(fused - o.recommend_dt(u)) <= 1e-12 * fused
    g.close()


def test_vacuum_light_wave_on_the_gpu():
    """the defining property once more, through the CUDA path: a plane wave moves at c (tests/test_maxwell_cpu.py)"""
    c = 2.0
    g = BoxSolver(2, 3, [8, 8], [0.0, 0.0], [1.0, 1.0], gamma=5.0 / 3.0, fields_enabled=True)
    g.set_maxwell(True, light_speed=c, chi=1.0, gamma=1.0)
    xyz = g.node_coords()
    u = np.zeros(g.shape)
    prim = np.zeros(xyz.shape[:-1] + (5,))
    prim[..., 0], prim[..., 4] = 1.0, 1.0
    cases.to_state(prim, 5.0 / 3.0, nc=13, species=0, u=u)
    u[:, 5 + 1, :] = np.cos(2 * np.pi * xyz[..., 0])
    u[:, 5 + 5, :] = np.cos(2 * np.pi * xyz[..., 0]) / c
    g.upload(0, u)
    g.solve(0.25)
    got = g.download(0)
    exact = np.cos(2 * np.pi * (xyz[..., 0] - c * 0.25))
    assert np.sqrt(np.mean((got[:, 6, :] - exact) ** 2)) < 1e-3
    g.close()


def test_two_fluid_langmuir_example_from_the_input_file():
    """examples/five-moment/two_fluid_langmuir.inp through the application (sources + Maxwell fluxes from the input keys)
    against the oracle driven the same way; the plasma oscillation is there: E_x builds up and the electron momentum comes back."""
    import os
    from warpii_b200 import App
    from test_gpu_input_file import oracle_for, oracle_initial_state, oracle_run
    here = os.path.dirname(os.path.abspath(__file__))
    text = open(os.path.join(here, "..", "examples", "five-moment", "two_fluid_langmuir.inp")).read()
    app = App(text)
    app.setup()
    steps = app.run()
    s = app.solver.get_state_global()
    assert steps > 100 and np.isfinite(s).all()            # c = 5 on h = 1/8, degree 3: dt ~ 1e-3
    assert len(app.frames) == 4 and abs(app.frames[-1][1] - 1.0) < 1e-12   # frame 0 fired in setup(), before run()
    o = oracle_for(app)
    o.set_sources(True, 0.01, 1.0, [1.0 / 25.0, -1.0])
    o.set_maxwell(True, light_speed=5.0, chi=1.0, gamma=1.0)
    u = oracle_initial_state(app, o)
    assert oracle_run(app, o, u) == steps
    rel = cases.rel_l2_per_component(s, u)
    assert (rel[[0, 4, 5, 6, 9, 10]] <= 1e-8).all(), rel   # densities, energies, electron momentum, E_x; 1280 steps (1e-10 per 100 steps)
    # the physics: E_x(t) = -(J0 / (eps0 omega_p)) sin(omega_p t) with omega_p^2 = sum_s (q/m)^2 rho / eps0 (cold plasma; the
    # thermal correction of the frequency is 3 % here)
    wp = np.sqrt((25.0 / 25.0 ** 2 + 1.0) / 0.01)
    want = 0.01 / (0.01 * wp) * abs(np.sin(wp * 1.0))
    assert abs(np.abs(s[:, 10, :]).max() - want) < 0.05 * want
    assert np.abs(s[:, 6, :]).max() < 0.01                 # the electrons are held in the oscillation
    app.close()
