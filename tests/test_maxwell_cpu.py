"""The perfectly hyperbolic Maxwell (PHM) fluxes of the field components: north_star kernel 4 / BASELINE config 5.

The reference only allocates [Ex,Ey,Ez,Bx,By,Bz,phi,psi] (five_moment.h:123-138) and evolves nothing there, so there is
no reference result to compare with (PARITY UNPINNED).  The oracle's definition (dgsem_oracle.cc::add_maxwell:
linear DGSEM flux on the fluid's nodes, Rusanov numerical flux) is therefore pinned by what defines the physics:
vacuum light-wave propagation at speed c with the scheme's order of accuracy, decay of a divergence error through the
cleaning potentials, the energy of the PHM system, and the exchange of energy with the fluids through the sources.
The GPU tier checks the CUDA path against this oracle (tests/test_gpu_maxwell.py)."""
import numpy as np
import pytest

import dgsem_cases as cases
import oracle
from oracle import Oracle


def vacuum(o, u, rho=1.0, p=1.0):
    """a fluid at rest (the operator needs one species); sources stay off, so it does not feel the fields"""
    prim = np.zeros(o.node_coords().shape[:-1] + (5,))
    prim[..., 0], prim[..., 4] = rho, p
    return cases.to_state(prim, o.gamma, nc=o.nc, species=0, u=u)


def phm_energy(o, u, c):
    """W = 1/2 int (|E|^2 + c^2 |B|^2 + c^2 phi^2 + psi^2): conserved by the PHM system in a periodic vacuum"""
    F = u[:, 5 * o.nsp:, :]
    dens = 0.5 * (F[:, 0] ** 2 + F[:, 1] ** 2 + F[:, 2] ** 2 + c * c * (F[:, 3] ** 2 + F[:, 4] ** 2 + F[:, 5] ** 2) + c * c * F[:, 6] ** 2
                  + F[:, 7] ** 2)
    w1, _ = None, None
    x, w = oracle.gll(o.p + 1)
    wN = w
    for _ in range(o.dim - 1):
        wN = np.multiply.outer(w, wN).reshape(-1)     # tensor weights, x fastest
    h = [(r - l) / n for l, r, n in zip(o.left, o.right, o.nx)]
    return float((dens * wN[None, :]).sum() * np.prod(h))


def light_wave(dim, p, n, c, t_end, direction=0):
    left, right = [0.0] * dim, [1.0] * dim
    o = Oracle(dim, p, [n] * dim, left, right, gamma=5.0 / 3.0, fields_enabled=True, threads=4)
    o.set_maxwell(True, light_speed=c, chi=1.0, gamma=1.0)
    xyz = o.node_coords()
    u = np.zeros(o.shape)
    vacuum(o, u)
    s = xyz[..., direction]
    k = 2 * np.pi
    # wave along e_d with E along e_{d+1} and B along e_{d+2}: E x B points along +e_d
    e_i, b_i = (direction + 1) % 3, (direction + 2) % 3
    u[:, 5 + e_i, :] = np.cos(k * s)
    u[:, 5 + 3 + b_i, :] = np.cos(k * s) / c
    steps = o.solve(u, t_end)
    exact = np.cos(k * (s - c * t_end))
    err = np.sqrt(np.mean((u[:, 5 + e_i, :] - exact) ** 2))
    errB = np.sqrt(np.mean((u[:, 5 + 3 + b_i, :] - exact / c) ** 2))
    other = [i for i in range(8) if i not in (e_i, 3 + b_i)]
    leak = np.abs(u[:, [5 + i for i in other], :]).max()
    return err, errB, leak, steps


@pytest.mark.parametrize("dim,direction", [(1, 0), (2, 0), (2, 1), (3, 2)])
def test_vacuum_light_wave_travels_at_c_with_the_scheme_order(dim, direction):
    c = 2.0
    n0 = 4 if dim < 3 else 3
    e1, b1, leak1, _ = light_wave(dim, 3, n0, c, 0.25, direction)
    e2, b2, leak2, _ = light_wave(dim, 3, 2 * n0, c, 0.25, direction)
    assert e2 < 2e-3 and abs(b2 - e2 / c) < 1e-12   # (B follows E: the wave is an eigenmode of the flux)
    assert e1 / e2 > 2 ** 3.3          # degree 3: order ~4 in space (time error of SSPRK2 at the CFL step is below it here)
    assert max(leak1, leak2) < 1e-12   # nothing else is excited: no spurious coupling between the components


def test_wave_speed_enters_recommend_dt():
    o = Oracle(2, 3, [4, 4], [0.0, 0.0], [1.0, 1.0], gamma=5.0 / 3.0, fields_enabled=True)
    u = np.zeros(o.shape)
    vacuum(o, u)
    dt0 = o.recommend_dt(u)
    o.set_maxwell(True, light_speed=50.0, chi=1.0, gamma=1.0)
    dt1 = o.recommend_dt(u)
    assert dt1 < dt0 / 10
    o.set_maxwell(True, light_speed=50.0, chi=2.0, gamma=1.0)      # the cleaning wave is the fastest one
    assert abs(o.recommend_dt(u) - dt1 / 2) < 1e-12 * dt1


def test_divergence_error_is_cleaned_and_energy_never_grows():
    """A magnetic field with a divergence (Bx = sin 2 pi x): with gamma > 0 the error is radiated into psi and damped by the
    upwinding; without cleaning it just sits there.  The PHM energy never increases (Rusanov flux), and is conserved to
    discretisation accuracy for the resolved wave."""
    c = 1.0
    def run(gam, t_end):
        o = Oracle(1, 3, [8], [0.0], [1.0], gamma=5.0 / 3.0, fields_enabled=True, threads=2)
        o.set_maxwell(True, light_speed=c, chi=1.0, gamma=gam)
        u = np.zeros(o.shape)
        vacuum(o, u)
        x = o.node_coords()[..., 0]
        u[:, 5 + 3, :] = np.sin(2 * np.pi * x)          # Bx(x): div B = 2 pi cos(2 pi x) != 0
        W = [phm_energy(o, u, c)]
        t = 0.0
        while t < t_end - 1e-12:
            dt = min(o.recommend_dt(u), t_end - t)
            o.ssprk2_step(u, dt, t)
            t += dt
            W.append(phm_energy(o, u, c))
        return o, u, np.array(W)
    o, u, W = run(1.0, 2.0)
    # fully discrete: SSPRK2 adds |R(i y)|^2 - 1 = y^4/4 per step to an oscillating mode; nothing more than that
    assert W[-1] < W[0] * (1 + 1e-5) and W[-1] > 0.99 * W[0]
    # semi-discrete: dW/dt = <F, dF/dt>_W <= 0 for ANY state (central volume terms cancel by the SBP property, the Rusanov
    # jump terms dissipate): random fields
    rng = np.random.default_rng(3)
    for dim, nx in ((1, [5]), (2, [3, 4]), (3, [2, 3, 2])):
        oo = Oracle(dim, 3, nx, [0.0] * dim, [1.0] * dim, gamma=5.0 / 3.0, fields_enabled=True, threads=2)
        oo.set_maxwell(True, light_speed=1.7, chi=1.3, gamma=0.6)
        v = np.zeros(oo.shape)
        vacuum(oo, v)
        v[:, 5:, :] = rng.standard_normal(v[:, 5:, :].shape)
        r, _ = oo.rhs(v)
        w_plus = phm_energy(oo, v + r * 1e-0, 1.7) - phm_energy(oo, v - r * 1e-0, 1.7)   # = 4 * <F, dF/dt>_W (W is quadratic)
        assert w_plus < 0 and w_plus / phm_energy(oo, v, 1.7) < -1e-3
        assert np.abs(r[:, :5, :]).max() < 1e-9          # the fluid at rest does not feel the fields (sources off)
    # the divergence error oscillates between Bx and psi instead of staying put: at a quarter period (t = 1/(4 gamma c)) it
    # has moved into psi entirely
    o, u, W = run(1.0, 0.25)
    assert np.abs(u[:, 5 + 3, :]).max() < 2e-3 and np.abs(u[:, 5 + 7, :]).max() > 0.99
    o, u, W = run(0.0, 0.25)                              # no cleaning: Bx is a steady state
    assert np.abs(u[:, 5 + 3, :] - np.sin(2 * np.pi * o.node_coords()[..., 0])).max() < 1e-12


def test_two_fluid_total_energy_with_fields():
    """Fluids + fields + sources: the electric work leaves the field energy and enters the fluids' total energy, so
    sum_s int E_s + eps0 W_EM only changes by the (dissipative, small) numerical fluxes: plasma oscillation in 2-D."""
    gamma, c, eps0 = 5.0 / 3.0, 3.0, 1.0
    o = Oracle(2, 3, [6, 6], [0.0, 0.0], [1.0, 1.0], gamma=gamma, n_species=2, fields_enabled=True, threads=4)
    qm = [1.0 / 25.0, -1.0]
    o.set_sources(True, epsilon0=eps0, chi=1.0, charge_over_mass=qm)
    o.set_maxwell(True, light_speed=c, chi=1.0, gamma=1.0)
    xyz = o.node_coords()
    s = np.sin(2 * np.pi * xyz[..., 0]) * np.cos(2 * np.pi * xyz[..., 1])
    u = np.zeros(o.shape)
    for sp, (rho0, vx) in enumerate([(25.0, 0.0), (1.0, 0.05)]):
        prim = np.zeros(xyz.shape[:-1] + (5,))
        prim[..., 0] = rho0            # charge neutral: q/m * rho sums to zero
        prim[..., 1] = vx * s          # the electrons are pushed: a Langmuir oscillation starts
        prim[..., 4] = 1.0
        cases.to_state(prim, gamma, nc=o.nc, species=sp, u=u)
    u[:, 10 + 5, :] = 0.1             # a uniform Bz, so that the magnetic force is active too

    def total(u):
        fl = sum(o.global_integral(u, sp)[4] for sp in range(2))
        return fl + eps0 * phm_energy(o, u, c), fl
    e0, f0 = total(u)
    t = 0.0
    moved, drift = 0.0, 0.0
    for _ in range(150):
        dt = o.recommend_dt(u)
        o.ssprk2_step(u, dt, t)
        t += dt
        now, fl = total(u)
        moved = max(moved, abs(fl - f0))
        drift = max(drift, abs(now - e0))
    assert moved > 1e-6                        # energy did move between fluids and field (it sloshes back and forth) ...
    assert drift < 2e-3 * moved                # ... while the total stayed put


def test_input_keys_for_the_field_system():
    """five_moment_maxwell / light_speed / phm_gamma of the application (warpii_b200/host/five_moment_app.hpp): parsing needs
    no GPU; the example input parses with them."""
    import os
    from warpii_b200 import App, WarpiiGpuError
    here = os.path.dirname(os.path.abspath(__file__))
    text = open(os.path.join(here, "..", "examples", "five-moment", "two_fluid_langmuir.inp")).read()
    app = App(text)
    assert app.fields_enabled and app.n_species == 2 and app.nx == [8, 4] and app.t_end == 1.0
    assert app.species(0)["mass"] == 25.0 and app.species(1)["charge"] == -1.0
    with pytest.raises(WarpiiGpuError, match="needs the field components"):
        App("set five_moment_maxwell = true\nset fields_enabled = false")
    with pytest.raises(WarpiiGpuError, match="outside the allowed range"):
        App("set n_species = 2\nset five_moment_maxwell = true\nset light_speed = 0")
    with pytest.raises(WarpiiGpuError, match="outside the allowed range"):
        App("set n_species = 2\nset five_moment_maxwell = true\nset phm_gamma = -1")
