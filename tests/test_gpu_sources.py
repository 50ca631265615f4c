"""GPU parity of the two-fluid source terms (north_star kernel 4) against the oracle's definition, plus physics checks on
the device.  The reference has no such terms: parity is GPU-vs-oracle only ('parity unpinned' upstream)."""
import numpy as np
import pytest

import oracle
import dgsem_cases as cases
from test_gpu_parity import RHS_TOL, STEPS_TOL, make_pair
from test_input_file_cpu import read_input
from warpii_b200 import App

pytestmark = pytest.mark.gpu


def plasma_state(o, rng):
    xyz = o.node_coords()
    k = 2 * np.pi
    s = np.sin(k * xyz[..., 0]) if o.dim == 1 else np.sin(k * xyz[..., 0]) * np.cos(k * xyz[..., 1])
    u = np.zeros(o.shape)
    for sp, (rho0, vel, p0) in enumerate([(25.0, (0.05, -0.02, 0.01), 0.8), (1.0, (-0.4, 0.3, 0.2), 0.6)]):
        prim = np.zeros(xyz.shape[:-1] + (5,))
        prim[..., 0] = rho0 * (1 + 0.1 * s)
        for d in range(3):
            prim[..., 1 + d] = vel[d] * (1 + 0.2 * s)
        prim[..., 4] = p0 * (1 + 0.05 * s)
        u[:, 5 * sp:5 * sp + 5, :] = oracle.primitive_to_conserved(prim, o.gamma).transpose(0, 2, 1)
    amp = rng.normal(size=8)
    for c in range(8):
        u[:, 10 + c, :] = amp[c] * (1 + 0.3 * s)
    return np.ascontiguousarray(u)


@pytest.mark.parametrize("dim,p,nx", [(1, 3, [12]), (2, 3, [8, 6]), (2, 2, [5, 7]), (3, 2, [4, 3, 3])])
def test_rhs_with_sources(dim, p, nx):
    o, g = make_pair(dim, p, nx, [0.0] * dim, [1.0] * dim, gamma=5.0 / 3.0, n_species=2, fields=True)
    qm = [1.0 / 25.0, -1.0]
    o.set_sources(True, 0.7, 1.3, qm)
    g.set_sources(True, 0.7, 1.3, qm)
    u = plasma_state(o, np.random.default_rng(3))
    g.upload_global(0, u)
    g.rhs(1, 0)
    got = g.download_global(1)
    want, _ = o.rhs(u)
    err = cases.rel_l2_per_component(got, want)
    active = [c for c in range(18) if np.abs(want[:, c]).max() > 0]
    assert set(active) >= {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 16}
    assert (err[active] <= RHS_TOL).all(), err
    assert np.all(got[:, [13, 14, 15, 17]] == 0.0)                   # B and psi have no source
    # switching the sources off again restores the reference's operator bit for bit
    g.set_sources(False)
    o.set_sources(False)
    g.rhs(1, 0)
    want0, _ = o.rhs(u)
    got0 = g.download_global(1)
    assert np.all(got0[:, 10:] == 0.0) and (cases.rel_l2_per_component(got0, want0)[:10] <= RHS_TOL).all()
    g.close()


def test_steps_with_sources_follow_the_oracle():
    o, g = make_pair(2, 3, [6, 6], [0.0, 0.0], [1.0, 1.0], gamma=5.0 / 3.0, n_species=2, fields=True)
    qm = [1.0 / 25.0, -1.0]
    o.set_sources(True, 1.0, 0.5, qm)
    g.set_sources(True, 1.0, 0.5, qm)
    u = plasma_state(o, np.random.default_rng(4))
    g.set_state_global(u)
    dt = 2e-4
    steps = g.solve(40 * dt, fixed_dt=dt)
    assert steps == 40 == o.solve(u, 40 * dt, fixed_dt=dt)
    err = cases.rel_l2_per_component(g.get_state_global(), u)
    assert (err <= STEPS_TOL).all(), err
    g.close()


def test_plasma_oscillation_and_energy_exchange_on_the_device():
    """Cold electrons against heavy ions: the momentum reverses after half a plasma period (omega_p = 1) and the sum of
    fluid and electrostatic energy is conserved to the accuracy of the time integration."""
    o, g = make_pair(1, 3, [16], [0.0], [1.0], gamma=5.0 / 3.0, n_species=2, fields=True)
    g.set_sources(True, 1.0, 0.0, [1.0e-6, -1.0])
    x = o.node_coords()[..., 0]
    ue, pe = 1e-3 * np.sin(2 * np.pi * x), 1e-6
    u = np.zeros(o.shape)
    u[:, 0, :] = 1.0e6
    u[:, 4, :] = pe / (o.gamma - 1)
    u[:, 5, :] = 1.0
    u[:, 6, :] = ue
    u[:, 9, :] = 0.5 * ue ** 2 + pe / (o.gamma - 1)
    g.set_state_global(u)

    def total_energy():
        s = g.get_state_global()
        fluid = g.global_integral(0, 0)[4] + g.global_integral(0, 1)[4]
        w = np.zeros(o.shape)
        w[:, 0, :] = 0.5 * (s[:, 10:13, :] ** 2).sum(axis=1)      # eps0 |E|^2 / 2 with eps0 = 1
        return fluid + o.global_integral(w, 0)[0], s

    e0, _ = total_energy()
    g.solve(np.pi, fixed_dt=0.0025)
    e1, s = total_energy()
    assert np.abs(s[:, 6, :] + ue).max() < 2e-2 * 1e-3          # momentum reversed
    exchanged = 0.25 * 1e-6                                      # ~ initial kinetic energy of the wave, 1/2 * <ue^2>
    assert abs(e1 - e0) < 1e-3 * exchanged
    g.close()


def test_cyclotron_rotation():
    """Uniform B_z, no electric feedback (huge eps0): the momentum rotates with omega_c = (q/m) B."""
    o, g = make_pair(2, 2, [4, 4], [0.0, 0.0], [1.0, 1.0], gamma=5.0 / 3.0, n_species=1, fields=True)
    g.set_sources(True, 1e30, 0.0, [2.0])
    u = np.zeros(o.shape)
    u[:, 0, :] = 1.0
    u[:, 1, :] = 0.1
    u[:, 4, :] = 0.5 * 0.01 + 1.0 / (o.gamma - 1)
    u[:, 5 + 5, :] = 0.5                                          # B_z: omega_c = 2 * 0.5 = 1
    g.set_state_global(u)
    g.solve(np.pi / 2, fixed_dt=0.0025)
    s = g.get_state_global()
    assert np.abs(s[:, 1, :]).max() < 1e-5 and np.abs(s[:, 2, :] + 0.1).max() < 1e-5
    assert np.abs(s[:, 0, :] - 1.0).max() < 1e-13 and np.abs(s[:, 4, :] - u[:, 4, :]).max() < 1e-9
    g.close()


def test_two_fluid_input_file_with_sources():
    text = """
set n_dims = 1
set n_species = 2
set five_moment_sources = true
set t_end = 0.2
set fe_degree = 3
set write_output = false
set n_writeout_frames = 2
subsection geometry
  set nx = 8
end
subsection Species_1
  set name = ion
  set charge = 1.0
  set mass = 1836.0
  subsection InitialCondition
    set Function expression = 1836.0; 0; 0; 0; 1e-3
  end
end
subsection Species_2
  set name = electron
  set charge = -1.0
  subsection InitialCondition
    set Function constants = pi=3.14
    set Function expression = 1.0; 1e-2*sin(2*pi*x); 0; 0; 1e-3
  end
end
"""
    app = App(text)
    app.setup()
    steps = app.run()
    s = app.solver.get_state_global()
    assert steps > 0 and np.isfinite(s).all()
    assert np.abs(s[:, 10, :]).max() > 1e-4                       # the electron current has built up an E_x
    # same run on the oracle
    from test_gpu_input_file import oracle_for, oracle_initial_state, oracle_run
    o = oracle_for(app)
    o.set_sources(True, 1.0, 0.0, [1.0 / 1836.0, -1.0])
    u = oracle_initial_state(app, o)
    assert oracle_run(app, o, u) == steps
    assert (cases.rel_l2_per_component(s, u)[[0, 4, 5, 6, 9, 10]] <= STEPS_TOL).all()
    app.close()
