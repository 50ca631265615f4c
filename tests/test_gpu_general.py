"""GPU parity on general geometry (curved / unstructured elements): warpii_gpu_set_geometry +
stage_kernel_general against the oracle's orc_create_general on the same tables (run on the B200 box, -m gpu).

Tolerances as on the Cartesian path (north_star): one RHS relative L2 <= 1e-12 per component, <= 1e-10 after 100 steps,
recommend_dt 1e-13, conservation to round-off.
"""
import numpy as np
import pytest

import mesh_cases as mc
from oracle import GeneralOracle, Oracle
from warpii_b200 import BC_INFLOW, BC_OUTFLOW, BC_WALL, BoxSolver
from warpii_b200.capi import MeshSolver, mapped_metrics

pytestmark = pytest.mark.gpu

GAMMA = 1.4
RHS_TOL = 1e-12
STEPS_TOL = 1e-10
DT_TOL = 1e-13


def metrics(dim, p, mesh, xyz):
    return mapped_metrics(dim, p, xyz, mesh["face_neighbor"], mesh["neighbor_face"], mesh["bf_elem"], mesh["bf_side"])


def make_pair(dim, p, mesh, xyz, n_boundaries=0, bc=None, n_species=1, fields=False, threads=4):
    geo = metrics(dim, p, mesh, xyz)
    o = GeneralOracle(dim, p, mesh, geo, n_boundaries=n_boundaries, bc_kinds=bc, gamma=GAMMA, n_species=n_species,
                      fields_enabled=fields, threads=threads)
    g = MeshSolver(dim, p, mesh, geo, n_boundaries=n_boundaries, bc_kinds=bc, gamma=GAMMA, n_species=n_species,
                   fields_enabled=fields)
    assert g.shape == o.shape
    return o, g


def rel_l2(got, want):
    """Relative L2 per component; a component much smaller than the largest one is measured against 1e-3 of that."""
    norms = np.array([np.linalg.norm(want[:, c, :]) for c in range(want.shape[1])])
    floor = 1e-3 * norms.max()
    return np.array([np.linalg.norm(got[:, c, :] - want[:, c, :]) / max(norms[c], floor, 1e-300) for c in range(want.shape[1])])


def explain(got, want, mesh, alpha, dim, p):
    """Where the two differ: worst element, its blend factor and face codes, error on face nodes vs interior nodes."""
    Np = p + 1
    err = np.abs(got - want)
    e = int(np.unravel_index(err.argmax(), err.shape)[0])
    idx = np.indices((Np,) * dim).reshape(dim, -1)
    on_face = ((idx == 0) | (idx == Np - 1)).any(axis=0)
    per_elem = err.max(axis=(1, 2))
    bad = np.argsort(per_elem)[::-1][:5]
    nbf = mesh["neighbor_face"]
    return (f"worst element {e}: max err {err[e].max():.3e} (|want| {np.abs(want[e]).max():.3e}), alpha {alpha[e]}, "
            f"nbr {mesh['face_neighbor'][e].tolist()}, codes {None if nbf is None else nbf[e].tolist()}, "
            f"face-node err {err[e][:, on_face].max():.3e}, interior-node err {err[e][:, ~on_face].max() if (~on_face).any() else 0:.3e}; "
            f"worst five elements {bad.tolist()} errs {per_elem[bad]}; elements with err > 1e-9: {(per_elem > 1e-9).sum()} of {len(per_elem)}")


def check_rhs(o, g, u, mesh, dim, p, tol=RHS_TOL):
    g.upload(0, u)
    g.rhs(1, 0)
    got = g.download(1)
    want, bif_want = o.rhs(u)
    assert np.isfinite(got).all()
    err = rel_l2(got, want)
    assert (err <= tol).all(), f"relative L2 per component {err}; " + explain(got, want, mesh, o.alpha(u), dim, p)
    a_g, a_o = g.shock_indicator(0), o.alpha(u)
    assert np.allclose(a_g, a_o, rtol=1e-10, atol=1e-12)
    if g.n_boundaries > 0:
        bif_got = g.boundary_fluxes(1)
        assert np.abs(bif_got - bif_want).max() <= 1e-12 * max(np.abs(bif_want).max(), 1e-300), (bif_got, bif_want)
    return err


def check_dt_and_integrals(o, g, u):
    g.upload(0, u)
    dt_g, dt_o = g.recommend_dt(0), o.recommend_dt(u)
    assert abs(dt_g - dt_o) <= DT_TOL * dt_o, (dt_g, dt_o)
    for sp in range(o.nsp):
        ig, io = g.global_integral(0, sp), o.global_integral(u, sp)
        assert np.abs(ig - io).max() <= 1e-13 * np.abs(io).max()


# ---- Cartesian geometry through the general kernels: must agree with the Cartesian GPU path and both oracles ----------
IDENTITY_CASES = [
    # dim, p, nx, periodic
    (1, 2, [12], True), (1, 3, [9], False), (2, 3, [8, 6], True), (2, 2, [7, 5], False), (2, 4, [4, 5], True),
    (2, 5, [3, 4], False), (3, 2, [4, 3, 5], True), (3, 3, [4, 4, 3], False), (3, 4, [3, 2, 3], True), (2, 1, [9, 8], True),
]


@pytest.mark.parametrize("dim,p,nx,periodic", IDENTITY_CASES)
def test_identity_geometry_matches_cartesian_paths(dim, p, nx, periodic):
    left, right = [0.0, -1.0, 0.5][:dim], [2.0, 1.5, 2.0][:dim]
    per = [int(periodic)] * dim
    kinds = [BC_WALL, BC_OUTFLOW, BC_INFLOW]
    bc = None if periodic else np.array([[kinds[f % 3] for f in range(2 * dim)]])
    mesh, xyz = mc.mapped_box(dim, p, nx, left, right, per)
    o, g = make_pair(dim, p, mesh, xyz, 0 if periodic else 2 * dim, bc)
    box_o = Oracle(dim, p, nx, left, right, per, gamma=GAMMA, bc_kinds=bc, threads=4)
    box_g = BoxSolver(dim, p, nx, left, right, periodic=per, gamma=GAMMA, n_boundaries=None if periodic else 2 * dim, bc_kinds=bc)
    if not periodic:
        q_in = mc.to_conserved(np.array([1.1, 0.3, 0.1 if dim > 1 else 0.0, 0.0, 0.9]), GAMMA)
        for f in range(2 * dim):
            for s in (o, g, box_o, box_g):
                s.set_inflow(0, f, q_in)
    prim = mc.periodic_state(GAMMA, left, right, dim)(xyz)
    mc.add_kinks(prim)
    u = mc.state_from(prim, GAMMA)
    assert (o.alpha(u) > 0).any()
    check_rhs(o, g, u, mesh, dim, p)
    check_dt_and_integrals(o, g, u)
    want_box, _ = box_o.rhs(u)
    g.upload(0, u)
    g.rhs(1, 0)
    box_g.upload_global(0, u)
    box_g.rhs(1, 0)
    assert (rel_l2(g.download(1), want_box) <= RHS_TOL).all()
    assert (rel_l2(g.download(1), box_g.download_global(1)) <= RHS_TOL).all()
    g.close()
    box_g.close()


# ---- curved meshes ---------------------------------------------------------------------------------------------------------
CURVED_CASES = [
    # dim, p, nx, periodic, amplitude
    (2, 3, [8, 8], True, 0.05), (2, 2, [6, 7], True, 0.04), (2, 4, [4, 4], True, 0.05), (2, 5, [3, 3], True, 0.03),
    (2, 3, [6, 5], False, 0.05), (3, 2, [4, 4, 3], True, 0.03), (3, 3, [3, 4, 3], True, 0.03), (3, 3, [3, 3, 4], False, 0.03),
    (3, 4, [2, 3, 2], True, 0.02), (1, 3, [10], True, 0.05),
]


@pytest.mark.parametrize("dim,p,nx,periodic,amp", CURVED_CASES)
def test_one_rhs_on_curved_mesh(dim, p, nx, periodic, amp):
    left, right = [0.0] * dim, [1.0, 1.2, 0.9][:dim]
    per = [int(periodic)] * dim
    kinds = [BC_OUTFLOW, BC_WALL, BC_INFLOW]
    bc = None if periodic else np.array([[kinds[f % 3] for f in range(2 * dim)]])
    mesh, xyz = mc.mapped_box(dim, p, nx, left, right, per, mc.wavy(left, right, amp))
    o, g = make_pair(dim, p, mesh, xyz, 0 if periodic else 2 * dim, bc)
    if not periodic:
        q_in = mc.to_conserved(np.array([1.05, 0.35, -0.1 if dim > 1 else 0.0, 0.05, 1.0]), GAMMA)
        for f in range(2 * dim):
            o.set_inflow(0, f, q_in)
            g.set_inflow(0, f, q_in)
    ref = mc.box_node_coords(dim, p, nx, left, right)
    prim = mc.periodic_state(GAMMA, left, right, dim)(ref)
    u_smooth = mc.state_from(prim.copy(), GAMMA)
    check_rhs(o, g, u_smooth, mesh, dim, p)            # alpha = 0 everywhere: volume + faces only
    mc.add_kinks(prim)
    u = mc.state_from(prim, GAMMA)
    assert (o.alpha(u) > 0).any()
    check_rhs(o, g, u, mesh, dim, p)                   # with the subcell-FV blend active
    check_dt_and_integrals(o, g, u)
    g.close()


def test_free_stream_on_curved_mesh_gpu():
    dim, p, nx, left, right = 2, 3, [8, 8], [0.0, 0.0], [1.0, 1.0]
    mesh, xyz = mc.mapped_box(dim, p, nx, left, right, [1, 1], mc.wavy(left, right, 0.05))
    o, g = make_pair(dim, p, mesh, xyz)
    prim = np.zeros(xyz.shape[:-1] + (5,))
    prim[...] = [1.2, 0.7, -0.4, 0.2, 0.9]
    u = mc.state_from(prim, GAMMA)
    g.upload(0, u)
    g.rhs(1, 0)
    assert np.abs(g.download(1)).max() <= 5e-12
    g.close()


def test_rotated_box_with_boundaries():
    dim, p, nx, left, right = 2, 3, [6, 5], [0.0, 0.0], [1.0, 0.8]
    bc = np.array([[BC_WALL, BC_OUTFLOW, BC_INFLOW, BC_WALL]])
    rot, R = mc.rotation2d(0.61)
    mesh, xyz = mc.mapped_box(dim, p, nx, left, right, [0, 0], rot)
    o, g = make_pair(dim, p, mesh, xyz, 4, bc)
    q_in = mc.to_conserved(np.array([1.05, 0.3, 0.2, 0.05, 1.0]), GAMMA)
    o.set_inflow(0, 2, q_in)
    g.set_inflow(0, 2, q_in)
    prim = mc.smooth_state(GAMMA, dim)(xyz)
    mc.add_kinks(prim)
    check_rhs(o, g, mc.state_from(prim, GAMMA), mesh, dim, p)
    g.close()


@pytest.mark.parametrize("p,n,warp", [(3, 2, 0.0), (3, 3, 0.03), (2, 4, 0.03), (4, 2, 0.02)])
def test_unstructured_quads(p, n, warp):
    """Three blocks around a vertex, local frames rotated cell by cell: every pairing of local faces, both orientations."""
    verts, cells = mc.hexagon_blocks(n)
    mesh, xyz = mc.quad_mesh(verts, cells, p, mc.hexagon_boundary_id, warp=mc.swirl_warp(warp) if warp else None)
    bc = np.array([[BC_WALL, BC_INFLOW, BC_OUTFLOW]])
    o, g = make_pair(2, p, mesh, xyz, 3, bc)
    q_in = mc.to_conserved(np.array([1.1, 0.3, 0.2, 0.0, 1.0]), GAMMA)
    o.set_inflow(0, 1, q_in)
    g.set_inflow(0, 1, q_in)
    prim = mc.smooth_state(GAMMA, 2)(xyz)
    check_rhs(o, g, mc.state_from(prim.copy(), GAMMA), mesh, 2, p)
    mc.add_kinks(prim)
    u = mc.state_from(prim, GAMMA)
    check_rhs(o, g, u, mesh, 2, p)
    check_dt_and_integrals(o, g, u)
    g.close()


def test_two_species_with_fields_and_sources_on_curved_mesh():
    dim, p, nx, left, right = 2, 3, [6, 6], [0.0, 0.0], [1.0, 1.0]
    mesh, xyz = mc.mapped_box(dim, p, nx, left, right, [1, 1], mc.wavy(left, right, 0.04))
    o, g = make_pair(dim, p, mesh, xyz, n_species=2, fields=True)
    ref = mc.box_node_coords(dim, p, nx, left, right)
    prim = mc.periodic_state(GAMMA, left, right, dim)(ref)
    mc.add_kinks(prim)
    u = mc.state_from(prim, GAMMA, n_species=2, fields=True)
    rng = np.random.default_rng(7)
    u[:, 10:, :] = 0.1 * rng.standard_normal(u[:, 10:, :].shape)
    check_rhs(o, g, u, mesh, dim, p)
    qm = [1.0, -25.0]
    o.set_sources(True, 1.0, 1.1, qm)
    g.set_sources(True, 1.0, 1.1, qm)
    check_rhs(o, g, u, mesh, dim, p)
    g.close()


# ---- time stepping ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", ["curved2d", "hexagon", "curved3d"])
def test_100_steps_and_conservation(case):
    if case == "curved2d":
        dim, p, nx, left, right = 2, 3, [8, 8], [0.0, 0.0], [1.0, 1.0]
        mesh, xyz = mc.mapped_box(dim, p, nx, left, right, [1, 1], mc.wavy(left, right, 0.05))
        o, g = make_pair(dim, p, mesh, xyz)
        prim = mc.periodic_state(GAMMA, left, right, dim)(mc.box_node_coords(dim, p, nx, left, right))
    elif case == "curved3d":
        dim, p, nx, left, right = 3, 2, [4, 4, 4], [0.0] * 3, [1.0] * 3
        mesh, xyz = mc.mapped_box(dim, p, nx, left, right, [1, 1, 1], mc.wavy(left, right, 0.03))
        o, g = make_pair(dim, p, mesh, xyz)
        prim = mc.periodic_state(GAMMA, left, right, dim)(mc.box_node_coords(dim, p, nx, left, right))
    else:
        dim, p = 2, 3
        verts, cells = mc.hexagon_blocks(3)
        mesh, xyz = mc.quad_mesh(verts, cells, p, mc.hexagon_boundary_id, warp=mc.swirl_warp(0.03))
        bc = np.array([[BC_WALL, BC_WALL, BC_WALL]])
        o, g = make_pair(dim, p, mesh, xyz, 3, bc)
        prim = mc.smooth_state(GAMMA, 2)(xyz)
    u = mc.state_from(prim, GAMMA)
    g.upload(0, u)
    i0 = g.global_integral(0)
    bif = np.zeros(5 * o.n_boundaries) if o.n_boundaries else None
    t = 0.0
    for step in range(100):
        dt_o = o.recommend_dt(u)
        dt_g = g.recommend_dt(0)
        assert abs(dt_g - dt_o) <= 1e-11 * dt_o, (step, dt_g, dt_o)
        o.ssprk2_step(u, dt_o, t, bif)
        g.ssprk2_step(dt_o, t)
        t += dt_o
    got = g.download(0)
    err = rel_l2(got, u)
    assert (err <= STEPS_TOL).all(), err
    i1 = g.global_integral(0)
    if case == "curved2d":   # periodic, metric identities hold in 2D: mass, momentum and energy conserved to round-off
        assert np.abs(i1 - i0).max() <= 1e-12 * np.abs(i0).max(), (i0, i1)
    elif case == "curved3d":   # mass and energy exchange is conservative; momentum sees the 3D free-stream defect
        assert abs(i1[0] - i0[0]) <= 1e-12 * abs(i0[0]) and abs(i1[4] - i0[4]) <= 1e-12 * abs(i0[4]), (i0, i1)
    else:   # closed box of walls: mass is conserved, and the GPU's boundary-integrated fluxes match the oracle's
        assert abs(i1[0] - i0[0]) <= 1e-12 * abs(i0[0])
        assert np.abs(g.boundary_fluxes(0) - bif).max() <= 1e-10 * max(np.abs(bif).max(), 1e-300)
    g.close()


def test_device_resident_loop_on_curved_mesh():
    """warpii_gpu_advance_to (device clock, CUDA-graph batches) == host-driven steps, bit for bit, on general geometry."""
    dim, p, nx, left, right = 2, 3, [8, 8], [0.0, 0.0], [1.0, 1.0]
    mesh, xyz = mc.mapped_box(dim, p, nx, left, right, [1, 1], mc.wavy(left, right, 0.05))
    o, g = make_pair(dim, p, mesh, xyz)
    prim = mc.periodic_state(GAMMA, left, right, dim)(mc.box_node_coords(dim, p, nx, left, right))
    u = mc.state_from(prim, GAMMA)
    g.upload(0, u)
    t_end, steps = g.advance_to(0.0, 1e9, max_steps=24)
    a = g.download(0)
    g.upload(0, u)
    t = 0.0
    for _ in range(24):
        dt = g.recommend_dt(0)
        g.ssprk2_step(dt, t)
        t += dt
    assert steps == 24 and abs(t - t_end) <= 1e-15 * t
    assert np.array_equal(a, g.download(0))
    o_u = u.copy()
    o.solve(o_u, 1e9, max_steps=24)
    assert (rel_l2(a, o_u) <= STEPS_TOL).all()
    g.close()


# ---- GridType = Extension through the application layer ---------------------------------------------------------------
EXTENSION_INPUT = """
set Application = FiveMoment
set n_dims = 2
set t_end = 0.02
set fe_degree = 3
set n_boundaries = 3
set gas_gamma = 1.4
set n_writeout_frames = 2
subsection geometry
    set GridType = Extension
end
subsection Species_1
    subsection InitialCondition
        set VariablesType = Primitive
        set Function constants = pi=3.1415926535
        set Function expression = 1.0 + 0.2*sin(1.3*x + 0.5*y); 0.4 + 0.1*cos(0.9*y); -0.2 + 0.1*sin(1.1*x); 0.1*cos(0.8*x - 0.6*y); 1.0 + 0.1*cos(0.7*x - 1.2*y)
    end
    subsection BoundaryConditions
        set 0 = Wall
        set 1 = Inflow
        set 2 = Outflow
        subsection 1_Inflow
            set VariablesType = Primitive
            set Function expression = 1.1; 0.3; 0.2; 0.0; 1.0
        end
    end
end
"""


def test_extension_grid_through_the_application(tmp_path):
    """An input file with GridType = Extension + the triangulation an extension would populate -> the same run as the
    oracle started from the same initial condition on the same mesh; frames carry the real node coordinates."""
    from test_general_geometry_cpu import hexagon_face_ids
    from warpii_b200 import App
    import dgsem_cases as cases
    p = 3
    verts, cells = mc.hexagon_blocks(3)
    app = App.with_triangulation(EXTENSION_INPUT, verts, cells, hexagon_face_ids(verts, cells))
    app.set_output_dir(str(tmp_path))
    app.set_device_loop(False)
    app.setup()
    mesh, xyz = mc.quad_mesh(verts, cells, p, mc.hexagon_boundary_id)
    assert np.abs(app.solver.node_coords() - xyz).max() <= 1e-15
    bc = np.array([[BC_WALL, BC_INFLOW, BC_OUTFLOW]])
    geo = metrics(2, p, mesh, xyz)
    o = GeneralOracle(2, p, mesh, geo, n_boundaries=3, bc_kinds=bc, gamma=GAMMA, threads=4)
    o.set_inflow(0, 1, mc.to_conserved(np.array([1.1, 0.3, 0.2, 0.0, 1.0]), GAMMA))
    u = mc.state_from(mc.smooth_state(GAMMA, 2)(xyz), GAMMA)
    assert (rel_l2(app.solver.get_state(), u) <= 1e-14).all()   # the parsed initial condition is the same function
    steps = app.run()
    # the oracle follows the same time loop: recommend_dt every step, stops at the two frame times
    t, n = 0.0, 0
    for stop in (0.01, 0.02):
        while t < stop - 1e-12:
            dt = min(o.recommend_dt(u), stop - t)
            o.ssprk2_step(u, dt, t)
            t += dt
            n += 1
    assert steps == n
    assert (rel_l2(app.solver.get_state(), u) <= STEPS_TOL).all()
    assert [f for f, _ in app.frames] == [1, 2]   # frame 0 was written by setup()
    frame = cases.read_vtu(str(tmp_path / "solution_002.vtu"))
    assert frame["n_points"] == xyz.shape[0] * xyz.shape[1]
    app.close()


def ffs_triangulation(rf=1):
    """The mesh of examples/five-moment/forward_facing_step/main.cc: 15 x 5 channel (times rf) minus the step cells."""
    Lx, Ly, nx, ny = 3.0, 1.0, 15 * rf, 5 * rf
    dx, dy = Lx / 15, Ly / 5
    hx, hy = Lx / nx, Ly / ny
    vid, verts, cells = {}, [], []

    def vertex(i, j):
        if (i, j) not in vid:
            vid[(i, j)] = len(verts)
            verts.append([i * hx, j * hy])
        return vid[(i, j)]

    for j in range(ny):
        for i in range(nx):
            if (i + 0.5) * hx > 3 * dx and (j + 0.5) * hy < dy:
                continue
            cells.append([vertex(i, j), vertex(i + 1, j), vertex(i, j + 1), vertex(i + 1, j + 1)])
    verts, cells = np.array(verts), np.array(cells)

    def bid(mid):
        tol = 1e-8
        if abs(mid[0]) < tol: return 0
        if abs(mid[1]) < tol: return 1
        if abs(mid[0] - 3 * dx) < tol and mid[1] < dy: return 2
        if abs(mid[1] - dy) < tol and mid[0] > 3 * dx: return 3
        if abs(mid[0] - Lx) < tol: return 4
        return 5
    return verts, cells, bid


def test_forward_facing_step_example(tmp_path):
    """The reference's extension example on the GPU path: the executable runs from the example input; the same mesh and
    input through the library match the oracle over the first steps of the Mach 3 start-up (shocks, alpha > 0)."""
    import os
    import subprocess
    from warpii_b200 import App
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "examples", "five-moment", "mach3_step.inp")).read()
    import re
    def setting(txt, key, value):
        new, n = re.subn(r"set %s\s*=.*" % key, "set %s = %s" % (key, value), txt)
        assert n == 1, key
        return new
    text = setting(setting(setting(text, "t_end", "0.01"), "n_writeout_frames", "1"), "RefinementFactor", "1")
    inp = tmp_path / "ffs.inp"
    inp.write_text(text)
    exe = os.path.join(root, "warpii_b200", "bin", "forward_facing_step")
    out = subprocess.run([exe, str(inp)], cwd=tmp_path, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "frame 1  t = 0.01" in out.stdout and (tmp_path / "FiveMoment__ffs" / "solution_001.vtu").exists()
    steps_exe = int(out.stdout.split("steps = ")[1].split()[0])

    p = 4
    verts, cells, bid = ffs_triangulation(1)
    ids = np.array([[bid(0.5 * (verts[c[a]] + verts[c[b]])) for (a, b) in mc._FACE_VERTS] for c in cells], dtype=np.int32)
    # (the array-fed extension of the C API declares no entries of its own)
    app = App.with_triangulation(re.sub(r"set RefinementFactor\s*=.*", "", text), verts, cells, ids)
    app.set_device_loop(False)
    app.setup()
    mesh, xyz = mc.quad_mesh(verts, cells, p, bid)
    bc = np.array([[BC_INFLOW, BC_WALL, BC_WALL, BC_WALL, BC_OUTFLOW, BC_WALL]])
    o = GeneralOracle(2, p, mesh, metrics(2, p, mesh, xyz), n_boundaries=6, bc_kinds=bc, gamma=1.4, threads=8)
    q_in = mc.to_conserved(np.array([1.4, 3.0, 0.0, 0.0, 1.0]), 1.4)
    o.set_inflow(0, 0, q_in)
    prim = np.zeros(xyz.shape[:-1] + (5,))
    prim[...] = [1.4, 3.0, 0.0, 0.0, 1.0]
    u = mc.state_from(prim, 1.4)
    steps = app.run()
    n = o.solve(u, 0.01)
    assert steps == n == steps_exe
    assert (o.alpha(u) > 0).any()
    got = app.solver.get_state()
    assert np.isfinite(got).all() and got[:, 0, :].min() > 0
    err = rel_l2(got, u)
    assert (err <= 1e-8).all(), err   # a shock run: alpha switches amplify round-off (same bar as the Sod run)
    app.close()


def test_mapped_box_solver_of_the_host_layer():
    """warpii_mapped_box_solver_create: the C++ host layer builds connectivity, support points and metric terms itself
    (GeneralMesh::mapped_box + build_mapped_metrics) and runs the SSPRK2 loop on curved elements."""
    dim, p, nx, left, right = 2, 3, [8, 6], [0.0, 0.0], [1.0, 1.2]
    warp = mc.wavy(left, right, 0.04)
    bc = [[BC_WALL, BC_OUTFLOW, BC_WALL, BC_WALL]]
    g = BoxSolver.mapped(dim, p, nx, left, right, lambda x: warp(x[None, :])[0], periodic=[0, 1], gamma=GAMMA, n_boundaries=4,
                         bc_kinds=bc)
    mesh, xyz = mc.mapped_box(dim, p, nx, left, right, [0, 1], warp)
    assert np.abs(g.node_coords() - xyz[g.l2g]).max() <= 1e-14
    o = GeneralOracle(dim, p, mesh, metrics(dim, p, mesh, xyz), n_boundaries=4, bc_kinds=np.array(bc), gamma=GAMMA, threads=4)
    prim = mc.periodic_state(GAMMA, left, right, dim)(mc.box_node_coords(dim, p, nx, left, right))
    mc.add_kinks(prim)
    u = mc.state_from(prim, GAMMA)
    g.set_state_global(u)
    steps = g.solve(0.01)
    n = o.solve(u, 0.01)
    assert steps == n and steps > 5
    assert (rel_l2(g.get_state_global(), u) <= STEPS_TOL).all()
    g.close()
