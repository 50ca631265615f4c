"""GPU tests of the widened rows of SURVEY section 8(f): space/time-dependent inflow and the input-file front end.

The oracle evaluates the inflow Function at every boundary quadrature point with the stage time, as the reference does
(fluid_flux_es_dgsem_operator.h:139-144, 381-384); the product tabulates it on the host and uploads the table through
warpii_gpu_set_inflow_table.  The App runs the reference's own test inputs (tests/golden/inputs) end to end and is
compared with the oracle started from the same parsed initial condition.  Tolerances: north_star's 1e-12 per RHS,
1e-10 after many steps.
"""
import math
import os

import numpy as np
import pytest

import oracle
from oracle import Oracle
import dgsem_cases as cases
from test_gpu_parity import RHS_TOL, STEPS_TOL, check_rhs, make_pair
from test_input_file_cpu import read_input
from warpii_b200 import App, BC_INFLOW, BC_OUTFLOW, BC_WALL

pytestmark = pytest.mark.gpu


def jet(gamma, pulse=0.0):
    """Conserved inflow state with a jet profile in y (and z), optionally pulsing in time."""
    def fn(x, t):
        r2 = sum((x[d] - 0.5) ** 2 for d in range(1, len(x)))
        rho = 1.4 * (1.0 + 0.2 * math.exp(-r2 / 0.04) * (math.cos(pulse * t) if pulse else 1.0))
        ux = 3.0 + 0.1 * math.sin(3.0 * x[-1])
        return [rho, rho * ux, 0.0, 0.0, 0.5 * rho * ux * ux + 1.0 / (gamma - 1)]
    return fn


def uniform(xyz):
    """Primitive Mach-3 free stream of the forward-facing-step example."""
    return np.tile([1.4, 3.0, 0.0, 0.0, 1.0], xyz.shape[:-1] + (1,))


def tabulate(g, fn, boundary_id, t=0.0):
    xyz, ids = g.boundary_points()
    table = np.zeros(xyz.shape[:2] + (5,))
    for f in np.nonzero(ids == boundary_id)[0]:
        for q in range(xyz.shape[1]):
            table[f, q] = fn(xyz[f, q], t)
    return table, ids


def test_boundary_points_match_the_oracle_order_independent():
    """The product's quadrature points are the oracle's, face by face (matched by coordinates)."""
    bc = [[BC_INFLOW, BC_OUTFLOW, BC_WALL, BC_WALL]]
    o, g = make_pair(2, 3, [6, 4], [0.0, 0.0], [1.5, 1.0], periodic=[0, 0], gamma=1.4, bc=bc)
    xyz, ids = g.boundary_points()
    assert xyz.shape == (2 * 6 + 2 * 4, 5, 2)
    assert sorted(set(ids.tolist())) == [0, 1, 2, 3]
    assert np.allclose(xyz[ids == 0][:, :, 0], 0.0) and np.allclose(xyz[ids == 1][:, :, 0], 1.5)
    assert np.allclose(xyz[ids == 2][:, :, 1], 0.0) and np.allclose(xyz[ids == 3][:, :, 1], 1.0)
    seen = []
    o.set_inflow_function(0, 0, lambda x, t: (seen.append(tuple(x)), [1.4, 4.2, 0, 0, 8.8])[1])
    o.rhs(o.project(uniform))
    mine = {tuple(np.round(p, 13)) for p in xyz[ids == 0].reshape(-1, 2)}
    theirs = {tuple(np.round(p, 13)) for p in seen}
    assert mine == theirs and len(mine) == 4 * 5
    g.close()


@pytest.mark.parametrize("dim,p,nx,right", [(2, 3, [6, 5], [1.5, 1.0]), (2, 2, [4, 7], [1.0, 1.0]), (3, 2, [3, 4, 3], [1.0, 1.0, 1.0])])
def test_space_dependent_inflow_rhs(dim, p, nx, right):
    gamma = 1.4
    bc = [[BC_INFLOW, BC_OUTFLOW] + [BC_WALL] * (2 * dim - 2)]
    o, g = make_pair(dim, p, nx, [0.0] * dim, right, periodic=[0] * dim, gamma=gamma, bc=bc)
    fn = jet(gamma)
    o.set_inflow_function(0, 0, fn)
    g.set_inflow_function(0, 0, fn, time_dependent=False)
    u = o.project(lambda xyz: uniform(xyz) * (1 + 0.05 * np.sin(2 * xyz[..., :1])))
    # the host layer tabulates the function on the first stage; rhs() goes straight to the ABI, so upload the table here
    g.set_inflow_table(0, tabulate(g, fn, 0)[0])
    check_rhs(o, g, u)
    # ... and a short run through the host layer's own tabulation (once, then the device-resident loop) follows the oracle
    g2 = make_pair(dim, p, nx, [0.0] * dim, right, periodic=[0] * dim, gamma=gamma, bc=bc)[1]
    g2.set_inflow_function(0, 0, fn, time_dependent=False)
    g2.set_state_global(u)
    t_end = 4 * g2.recommend_dt()
    steps = g2.solve(t_end)
    assert steps == o.solve(u, t_end) and steps >= 4
    assert (cases.rel_l2_per_component(g2.get_state_global(), u)[[0, 1, 4]] < STEPS_TOL).all()
    g.close()
    g2.close()


def test_constant_inflow_after_table_overrides_rows():
    gamma = 1.4
    bc = [[BC_INFLOW, BC_INFLOW, BC_WALL, BC_WALL]]
    o, g = make_pair(2, 2, [4, 4], [0.0, 0.0], [1.0, 1.0], periodic=[0, 0], gamma=gamma, bc=bc)
    fn = jet(gamma)
    q1 = oracle.primitive_to_conserved([1.2, -2.5, 0.0, 0.0, 0.9], gamma)
    o.set_inflow_function(0, 0, fn)
    o.set_inflow(0, 1, q1)
    table, ids = tabulate(g, fn, 0)
    table[ids != 0] = 7.0             # garbage in the rows of every other face
    g.set_inflow_table(0, table)
    g.set_inflow(0, 1, q1)            # a constant state set after the table overwrites the rows of its boundary
    u = o.project(uniform)
    check_rhs(o, g, u)
    g.close()


def test_time_dependent_inflow_steps_match_oracle():
    """SSPRK2 stages see the inflow at t and t + dt (rk.h:97-106 + set_time, :139-144): host-driven loop vs oracle."""
    gamma = 1.4
    bc = [[BC_INFLOW, BC_OUTFLOW, BC_WALL, BC_WALL]]
    o, g = make_pair(2, 3, [6, 4], [0.0, 0.0], [1.5, 1.0], periodic=[0, 0], gamma=gamma, bc=bc, threads=1)
    fn = jet(gamma, pulse=300.0)
    o.set_inflow_function(0, 0, fn)
    g.set_inflow_function(0, 0, fn, time_dependent=True)
    u = o.project(uniform)
    g.set_state_global(u)
    t_end = 0.03
    bif_o = np.zeros(5 * 4)
    steps_g = g.solve(t_end)
    steps_o = o.solve(u, t_end, bif=bif_o)
    assert steps_g == steps_o and steps_g >= 10
    got = g.get_state_global()
    err = cases.rel_l2_per_component(got, u)
    assert (err[[0, 1, 4]] < STEPS_TOL).all(), err
    # the transverse momentum is tiny here: measure its error against the momentum vector, not against itself
    assert np.linalg.norm(got[:, 2] - u[:, 2]) < STEPS_TOL * np.linalg.norm(u[:, 1])
    assert np.allclose(g.boundary_fluxes(0), bif_o, rtol=1e-11, atol=1e-13)
    # the pulse must actually have been seen: a frozen-in-time inflow gives a different answer
    o2, g2 = make_pair(2, 3, [6, 4], [0.0, 0.0], [1.5, 1.0], periodic=[0, 0], gamma=gamma, bc=bc, threads=1)
    g2.set_inflow_function(0, 0, jet(gamma, pulse=0.0), time_dependent=False)
    g2.set_state_global(o2.project(uniform))
    g2.solve(t_end)
    assert cases.rel_l2_per_component(g2.get_state_global(), u)[0] > 1e-6
    g.close()
    g2.close()


def oracle_for(app, threads=4):
    bc = None
    if app.n_boundaries:
        bc = [app.species(s)["bc_kinds"] + [BC_WALL] * (2 * app.n_dims - app.n_boundaries) for s in range(app.n_species)]
    return Oracle(app.n_dims, app.fe_degree, app.nx, app.left, app.right, periodic=[int(p) for p in app.periodic],
                  gamma=app.gas_gamma, n_species=app.n_species, fields_enabled=app.fields_enabled, bc_kinds=bc, threads=threads)


def oracle_initial_state(app, o):
    u = np.zeros(o.shape)
    xyz = o.node_coords()
    for s in range(app.n_species):
        q, _ = app.eval_function(s, xyz.reshape(-1, app.n_dims))
        u[:, 5 * s:5 * s + 5, :] = q.reshape(o.n_elems, o.NN, 5).transpose(0, 2, 1)
    return u


def oracle_run(app, o, u):
    """FiveMomentApp::run on the oracle: advance() with the writeout callback of five_moment.h:233-243, which clips dt at
    every frame time.  Returns the number of steps."""
    steps = [0]

    def step(t, dt):
        o.ssprk2_step(u, dt, t)
        steps[0] += 1
        return True

    interval = app.t_end / app.n_writeout_frames
    oracle.advance(step, app.t_end, lambda: o.recommend_dt(u), [(interval, lambda t: None, False, True)])
    return steps[0]


def test_reference_freestream_1d_from_the_input_file():
    """InputTest.FreeStream1D (test/input_test.cc:18-66) run the way the reference runs it: from the input text."""
    from test_oracle_golden import _l2_error_density
    errs = []
    for nx in (20, 30):
        app = App(read_input("freestream_1d.inp") + f"subsection geometry\n set nx = {nx}\n end")
        app.setup()
        steps = app.run()
        o = oracle_for(app)
        u = oracle_initial_state(app, o)
        got = app.solver.get_state_global()
        errs.append(_l2_error_density(o, got, lambda x: 1 + 0.6 * np.sin(2 * np.pi * (x - 0.04)), 2))
        assert oracle_run(app, o, u) == steps
        assert (cases.rel_l2_per_component(got, u)[[0, 1, 4]] < STEPS_TOL).all()
        # write_output = false: frames fire (0 in setup, 1..10 in run) but nothing is written
        assert [f for f, _ in app.frames] == list(range(1, 11)) and abs(app.frames[-1][1] - 0.04) < 1e-12
        app.close()
    assert abs(errs[1]) < 1e-4
    assert abs(errs[0] / errs[1] - 1.5 ** 3) < 1.0


def test_reference_sod_and_pseudo_2d_inputs_run():
    """InputTest.SodShocktube / FreeStreamPseudo2D / FreeStream2DDiagonal (test/input_test.cc:68-168): the reference only asks
    them to run; here they also have to follow the oracle."""
    app = App(read_input("sod_shocktube.inp"))
    app.setup()
    o = oracle_for(app)
    u = oracle_initial_state(app, o)
    assert np.array_equal(app.solver.get_state_global(), u)          # nodal interpolation of the parsed IC, bit for bit
    steps = app.run()
    got = app.solver.get_state_global()
    assert np.isfinite(got).all() and got[:, 0].min() > 0.05
    assert oracle_run(app, o, u) == steps
    err = cases.rel_l2_per_component(got, u)
    assert err[0] < 1e-8 and err[1] < 1e-8 and err[4] < 1e-8, err      # shock run: alpha switches are chaotic in the last bits
    app.close()

    for name, comps in (("freestream_pseudo_2d.inp", [0, 1, 4]), ("freestream_2d_diagonal_test.inp", [0, 1, 2, 4])):
        app = App(read_input(name))
        app.setup()
        ic = app.solver.global_integral(0)
        steps = app.run()
        assert steps > 10
        assert np.allclose(app.solver.global_integral(0), ic, rtol=0, atol=1e-13)
        o = oracle_for(app)
        u = oracle_initial_state(app, o)
        assert oracle_run(app, o, u) == steps
        assert (cases.rel_l2_per_component(app.solver.get_state_global(), u)[comps] < STEPS_TOL).all()
        app.close()


def test_device_loop_equals_host_driven_loop():
    states = []
    for device_loop in (True, False):
        app = App(read_input("freestream_2d_diagonal.inp"))
        app.set_device_loop(device_loop)
        app.setup()
        steps = app.run()
        states.append((steps, app.solver.get_state_global(), list(app.frames)))
        app.close()
    assert states[0][0] == states[1][0] > 5
    assert np.array_equal(states[0][1], states[1][1])
    assert states[0][2] == states[1][2]


def test_inflow_channel_input_with_pulsing_jet(tmp_path):
    """An input with a (y, t)-dependent inflow function: App (host-driven stages) vs the oracle with the same parsed function."""
    text = read_input("inflow_channel_2d.inp").replace("set write_output = false", "set write_output = true")
    app = App(text)
    app.set_output_dir(str(tmp_path))
    app.setup()
    o = oracle_for(app, threads=1)
    u = oracle_initial_state(app, o)
    o.set_inflow_function(0, 0, lambda x, t: app.eval_function(0, [x], t=t, boundary_id=0)[0][0])
    steps = app.run()
    assert oracle_run(app, o, u) == steps and steps > 10
    got = app.solver.get_state_global()
    assert (cases.rel_l2_per_component(got, u)[[0, 1, 4]] < STEPS_TOL).all()
    assert np.linalg.norm(got[:, 2] - u[:, 2]) < STEPS_TOL * np.linalg.norm(u[:, 1])
    # frames: 0 from setup, 1..4 from run; raw files hold the state in device order
    assert [f for f, _ in app.frames] == [1, 2, 3, 4]
    assert np.allclose([t for _, t in app.frames], [0.005, 0.01, 0.015, 0.02], rtol=0, atol=1e-12)
    names = sorted(os.listdir(tmp_path))
    assert names == ["frames.txt"] + [f"solution_{i:03d}.vtu" for i in range(5)]
    # the last frame holds the final state: names of five_moment.h:259-300, derived fields of postprocessor.h:33-62
    vtu = cases.read_vtu(tmp_path / "solution_004.vtu")
    state = app.solver.get_state()                                    # device order, like the file
    n_el, NN = state.shape[0], state.shape[2]
    assert vtu["n_points"] == n_el * NN and vtu["n_cells"] == n_el * 9 and (vtu["types"] == 9).all()
    assert np.array_equal(vtu["neutral_density"], state[:, 0].reshape(-1))
    assert np.array_equal(vtu["neutral_x_momentum"], state[:, 1].reshape(-1))
    assert np.array_equal(vtu["neutral_energy"], state[:, 4].reshape(-1))
    rho, mx, my, E = (state[:, c].reshape(-1) for c in (0, 1, 2, 4))
    p = (app.gas_gamma - 1) * (E - (mx ** 2 + my ** 2) / (2 * rho))
    np.testing.assert_allclose(vtu["pressure"], p, rtol=1e-13)
    np.testing.assert_allclose(vtu["x_velocity"], mx / rho, rtol=1e-15)
    np.testing.assert_allclose(vtu["specific_entropy"], np.log(p) - app.gas_gamma * np.log(rho), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(vtu["speed_of_sound"], np.sqrt(app.gas_gamma * p / rho), rtol=1e-13)
    assert np.array_equal(vtu["Points"][:, :2], app.solver.node_coords().reshape(-1, 2)) and (vtu["owner"] == 0).all()
    # sub-cells are counter-clockwise quads of neighbouring GLL nodes of ONE element
    quad = vtu["connectivity"].reshape(-1, 4)
    assert (quad // NN == (quad[:, :1] // NN)).all() and (vtu["offsets"] == 4 * np.arange(1, len(quad) + 1)).all()
    xy = vtu["Points"][quad][:, :, :2]
    area2 = ((xy[:, 1, 0] - xy[:, 0, 0]) * (xy[:, 2, 1] - xy[:, 0, 1]) - (xy[:, 2, 0] - xy[:, 0, 0]) * (xy[:, 1, 1] - xy[:, 0, 1]))
    assert (area2 > 0).all() and abs(area2.sum() - 1.5 * 1.0) < 1e-12               # the sub-cells tile the 1.5 x 1 box
    lines = (tmp_path / "frames.txt").read_text().splitlines()
    assert len(lines) == 5 and lines[0].split()[0] == "0" and abs(float(lines[-1].split()[1]) - 0.02) < 1e-12
    app.close()


def test_cli_runs_an_input_file(tmp_path):
    import subprocess
    exe = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "warpii_b200", "bin", "warpii_gpu"))
    assert os.path.exists(exe), "warpii_gpu was not built (make -C warpii_b200)"
    inp = tmp_path / "fs1d.inp"
    inp.write_text(read_input("freestream_1d.inp").replace("set write_output = false", "set n_writeout_frames = 2")
                   + "subsection geometry\n set nx = 16\n end")
    r = subprocess.run([exe, "fs1d.inp"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert r.returncode == 0, r.stderr
    out_dir = tmp_path / "FiveMoment__fs1d"                                   # WorkDir = %A__%I, warpii.cc:139-149
    assert sorted(os.listdir(out_dir)) == ["frames.txt", "solution_000.vtu", "solution_001.vtu", "solution_002.vtu"]
    app = App(inp.read_text())
    app.setup()
    app.run()
    assert np.array_equal(cases.read_vtu(out_dir / "solution_002.vtu")["neutral_density"], app.solver.get_state()[:, 0].reshape(-1))
    app.close()
    r = subprocess.run([exe, "--setup-only", "fs1d.inp"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert r.returncode == 0 and "steps =" not in r.stdout
