"""Two-fluid source terms (north_star kernel 4) on the CPU oracle: new physics, not in the reference, so the
checks are self-consistency identities rather than reference goldens.  'Parity unpinned' by construction."""
import numpy as np

import oracle
from oracle import Oracle
from warpii_b200 import App, WarpiiGpuError
import pytest


def two_fluid_state(o, rng, uniform=True):
    """ion + electron fluids and 8 field components; uniform in space so that the flux divergence vanishes."""
    u = np.zeros(o.shape)
    prim = [[25.0, 0.03, -0.02, 0.01, 0.8], [1.0, -0.4, 0.3, 0.2, 0.6]]
    for s, w in enumerate(prim):
        u[:, 5 * s:5 * s + 5, :] = oracle.primitive_to_conserved(w, o.gamma)[None, :, None]
    u[:, 10:18, :] = rng.normal(size=8)[None, :, None] * (1.0 if uniform else 1.0)
    return u


def test_sources_vanish_when_disabled_and_field_rates_are_zero():
    o = Oracle(1, 2, [4], [0.0], [1.0], gamma=5.0 / 3.0, n_species=2, fields_enabled=True)
    u = two_fluid_state(o, np.random.default_rng(0))
    r, _ = o.rhs(u)
    assert np.abs(r[:, :10]).max() < 1e-12 and np.all(r[:, 10:] == 0.0)   # the reference's operator: fields untouched


def test_lorentz_force_current_and_energy_exchange():
    o = Oracle(2, 2, [3, 2], [0.0, 0.0], [1.0, 1.0], gamma=5.0 / 3.0, n_species=2, fields_enabled=True)
    qm = np.array([1.0 / 25.0, -1.0])
    eps0, chi = 0.7, 1.3
    o.set_sources(True, eps0, chi, qm)
    u = two_fluid_state(o, np.random.default_rng(1))
    r, _ = o.rhs(u)
    E, B = u[0, 10:13, 0], u[0, 13:16, 0]
    J = np.zeros(3)
    rho_c = 0.0
    for s in range(2):
        rho, m = u[0, 5 * s, 0], u[0, 5 * s + 1:5 * s + 4, 0]
        np.testing.assert_allclose(r[:, 5 * s, :], 0.0, atol=1e-12)                                     # no mass source
        np.testing.assert_allclose(r[0, 5 * s + 1:5 * s + 4, 0], qm[s] * (rho * E + np.cross(m, B)), rtol=1e-13, atol=1e-13)
        np.testing.assert_allclose(r[0, 5 * s + 4, 0], qm[s] * m.dot(E), rtol=1e-13, atol=1e-13)
        J += qm[s] * m
        rho_c += qm[s] * rho
    np.testing.assert_allclose(r[0, 10:13, 0], -J / eps0, rtol=1e-14)
    np.testing.assert_array_equal(r[:, 13:16, :], 0.0)                                                 # no curl terms here
    np.testing.assert_allclose(r[0, 16, 0], chi * rho_c / eps0, rtol=1e-14)
    np.testing.assert_array_equal(r[:, 17, :], 0.0)
    # energy exchange: what the fluids gain is what the field loses, d/dt (sum_s E_s + eps0 |E|^2 / 2) = 0 pointwise
    gain = r[:, 4, :] + r[:, 9, :] + eps0 * np.einsum("k,ekn->en", E, r[:, 10:13, :])
    assert np.abs(gain).max() < 1e-13
    # the magnetic force does no work
    o.set_sources(True, eps0, chi, qm)
    u2 = u.copy()
    u2[:, 10:13, :] = 0.0
    r2, _ = o.rhs(u2)
    assert np.abs(r2[:, 4, :]).max() < 1e-13 and np.abs(r2[:, 9, :]).max() < 1e-13


def test_cold_plasma_oscillation_frequency():
    """Electrostatic oscillation of a cold electron fluid against heavy ions: omega_p^2 = n q^2 / (m eps0) = 1."""
    o = Oracle(1, 2, [8], [0.0], [1.0], gamma=5.0 / 3.0, n_species=2, fields_enabled=True, threads=4)
    o.set_sources(True, 1.0, 0.0, [1.0e-6, -1.0])
    xyz = o.node_coords()[..., 0]
    u = np.zeros(o.shape)
    ue = 1e-3 * np.sin(2 * np.pi * xyz)
    pe = 1e-6
    u[:, 0, :] = 1.0e6                                    # ions: number density 1, mass 1e6, at rest
    u[:, 4, :] = pe / (o.gamma - 1)
    u[:, 5, :] = 1.0                                      # electrons: number density 1, mass 1
    u[:, 6, :] = ue
    u[:, 9, :] = 0.5 * ue ** 2 + pe / (o.gamma - 1)
    m0 = u[:, 6, :].copy()
    dt, t = 0.005, 0.0
    while t < np.pi - 1e-12:                              # half a plasma period: the electron momentum has reversed
        step = min(dt, np.pi - t)
        o.ssprk2_step(u, step, t)
        t += step
    np.testing.assert_allclose(u[:, 6, :], -m0, atol=2e-2 * np.abs(m0).max())
    assert np.abs(u[:, 10, :]).max() < 0.05 * np.abs(m0).max()     # E_x ~ sin(omega_p t) passes through zero


def test_input_keys_for_the_sources():
    text = """
set n_species = 2
set five_moment_sources = true
set epsilon0 = 2.5
set phm_chi = 1.5
subsection Species_1
  set charge = 1.0
  set mass = 25.0
end
subsection Species_2
  set charge = -1.0
end
"""
    app = App(text)
    assert app.fields_enabled
    with pytest.raises(WarpiiGpuError, match="needs the field components"):
        App("set five_moment_sources = true\nset fields_enabled = false")
    with pytest.raises(WarpiiGpuError, match="outside the allowed range"):
        App("set n_species = 2\nset epsilon0 = -1")
