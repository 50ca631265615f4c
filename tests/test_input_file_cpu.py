"""The input-file front end (warpii_b200/host/{parameter_file,expression,five_moment_app}.hpp) on the CPU: parsing needs no GPU.

What is checked: the reference's own test inputs and example inputs parse to the parameters the reference would read
(five_moment.h:99-198, species.cc:9-66, grid_descriptions.cc:27-49), ParsedFunction-style expressions evaluate like their
numpy transcription, Primitive -> conserved follows species_func.cc:9-30, and malformed inputs fail with a message.
"""
import math
import os

import numpy as np
import pytest

from warpii_b200 import App, WarpiiGpuError, BC_INFLOW, BC_OUTFLOW, BC_WALL

INPUTS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "inputs")


def read_input(name):
    with open(os.path.join(INPUTS, name)) as f:
        return f.read()


def eval_expr(expr, pts, t=0.0, constants="", dim=2):
    """Evaluate one expression through the product's parser (as conserved component 0 of an initial condition)."""
    geometry = {1: ("0", "1", "1"), 2: ("0,0", "1,1", "1,1"), 3: ("0,0,0", "1,1,1", "1,1,1")}[dim]
    text = f"""
set n_dims = {dim}
subsection geometry
  set left = {geometry[0]}
  set right = {geometry[1]}
  set nx = {geometry[2]}
end
set n_boundaries = 1
subsection Species_1
  subsection BoundaryConditions
    set 0 = Inflow
    subsection 0_Inflow
      set VariablesType = Conserved
      set Function constants = {constants}
      set Function expression = {expr}; 0; 0; 0; 0
    end
  end
end
"""
    app = App(text)
    q, time_dependent = app.eval_function(0, np.atleast_2d(pts), t=t, boundary_id=0)
    return q[:, 0], time_dependent


def test_reference_test_inputs_parse():
    a = App(read_input("freestream_1d.inp") + "subsection geometry\n set nx = 30\n end")   # the way input_test.cc:52-54 appends nx
    assert (a.n_dims, a.n_species, a.n_boundaries, a.fe_degree) == (1, 1, 0, 2)
    assert a.nx == [30] and a.left == [0.0] and a.right == [1.0] and a.periodic == [True]
    assert a.t_end == 0.04 and not a.write_output and not a.fields_enabled
    assert a.gas_gamma == 1.6666666666667 and a.n_writeout_frames == 10   # the declared defaults, five_moment.h:140-146

    s = App(read_input("sod_shocktube.inp"))
    assert (s.n_dims, s.n_boundaries, s.fe_degree, s.nx, s.periodic) == (1, 2, 4, [100], [False])
    assert s.species(0)["bc_kinds"] == [BC_OUTFLOW, BC_OUTFLOW]
    q, td = s.eval_function(0, [[0.25], [0.75]])
    g = s.gas_gamma
    assert not td
    np.testing.assert_array_equal(q, [[1.0, 0, 0, 0, 1.0 / (g - 1)], [0.1, 0, 0, 0, 0.125 / (g - 1)]])

    p = App(read_input("freestream_pseudo_2d.inp"))
    assert (p.n_dims, p.nx, p.right, p.periodic) == (2, [100, 2], [1.0, 0.02], [True, True])

    d = App(read_input("freestream_2d_diagonal.inp"))   # line continuations
    assert (d.n_dims, d.nx, d.fe_degree) == (2, [12, 12], 3)
    pts = np.random.default_rng(0).random((50, 2))
    q, _ = d.eval_function(0, pts)
    rho = 1 + 0.6 * np.sin(2 * math.pi * (pts[:, 0] + pts[:, 1]))   # "pi" is the exact constant whatever the input says
    np.testing.assert_allclose(q[:, 0], rho, rtol=0, atol=2e-16 * 4)
    np.testing.assert_allclose(q[:, 4], 0.5 * rho * 2.0 + 1.0 / (d.gas_gamma - 1), rtol=1e-15)


def test_species_and_boundary_conditions():
    a = App(read_input("inflow_channel_2d.inp"))
    assert (a.n_dims, a.n_boundaries, a.gas_gamma, a.periodic, a.n_writeout_frames) == (2, 4, 1.4, [False, False], 4)
    sp = a.species(0)
    assert sp == dict(name="neutral", charge=0.0, mass=1.0, bc_kinds=[BC_INFLOW, BC_OUTFLOW, BC_WALL, BC_WALL])
    pts = np.array([[0.0, 0.5], [0.0, 0.1]])
    q, time_dependent = a.eval_function(0, pts, t=0.03, boundary_id=0)
    assert time_dependent
    rho = 1.4 * (1 + 0.2 * np.exp(-((pts[:, 1] - 0.5) / 0.2) ** 2) * math.sin(40.0 * 0.03))
    np.testing.assert_allclose(q[:, 0], rho, rtol=1e-15)
    np.testing.assert_allclose(q[:, 1], 3.0 * rho, rtol=1e-15)
    np.testing.assert_allclose(q[:, 4], 0.5 * rho * 9.0 + 1.0 / 0.4, rtol=1e-15)
    _, ic_td = a.eval_function(0, pts)
    assert not ic_td
    with pytest.raises(WarpiiGpuError, match="no inflow function"):
        a.eval_function(0, pts, boundary_id=1)

    two = App("""
set n_species = 2
subsection Species_1
  set name = ion
  set charge = 1.0
  set mass = 25.0
end
subsection Species_2
  set name = electron
  set charge = -1.0
end
""")
    assert two.fields_enabled                           # "auto": enabled iff n_species >= 2 (five_moment.h:176-177)
    assert two.species(0)["name"] == "ion" and two.species(0)["mass"] == 25.0
    assert two.species(1) == dict(name="electron", charge=-1.0, mass=1.0, bc_kinds=[])
    assert not App("set n_species = 2\nset fields_enabled = false").fields_enabled
    assert App("set fields_enabled = true").fields_enabled


def test_expression_language():
    rng = np.random.default_rng(1)
    pts = rng.random((40, 2)) + 0.1
    x, y = pts[:, 0], pts[:, 1]
    cases = [
        ("x + y * 2 - 3 / 4", x + y * 2 - 0.75),
        ("-x^2", -(x ** 2)),                    # unary minus binds weaker than ^ (muparser)
        ("2^3^2 + 0*x", np.full_like(x, 512.0)),  # ^ is right associative
        ("(x - y) * -(x + y)", (x - y) * -(x + y)),
        ("sin(x)*cos(y) + tan(x) - exp(-y) + log(x) + sqrt(y) + abs(x - y)",
         np.sin(x) * np.cos(y) + np.tan(x) - np.exp(-y) + np.log(x) + np.sqrt(y) + np.abs(x - y)),
        ("tanh(4*(x-0.5)) + sinh(y) + cosh(y) + atan(x) + asin(y/2) + acos(x/2)",
         np.tanh(4 * (x - 0.5)) + np.sinh(y) + np.cosh(y) + np.arctan(x) + np.arcsin(y / 2) + np.arccos(x / 2)),
        ("pow(x, 3) + min(x, y) + max(x, y) + x^0.5", x ** 3 + np.minimum(x, y) + np.maximum(x, y) + x ** 0.5),
        ("if(x < 0.5, 1.0, 0.10)", np.where(x < 0.5, 1.0, 0.1)),
        ("if(x < 0.5 && y >= 0.4, 2, if(x > 0.8 || y == 7, 3, 4))",
         np.where((x < 0.5) & (y >= 0.4), 2.0, np.where(x > 0.8, 3.0, 4.0))),
        ("x > y ? x : y", np.maximum(x, y)),
        ("1e-3*x + 2.5E2*y + .5", 1e-3 * x + 250.0 * y + 0.5),
        ("floor(3*x) + ceil(3*y) + sign(x - y) + log10(x) + log2(y)",
         np.floor(3 * x) + np.ceil(3 * y) + np.sign(x - y) + np.log10(x) + np.log2(y)),
        # the rest of muparser's built-ins that deal.II's FunctionParser exposes
        ("erf(x) + erfc(y) + asinh(x) + acosh(1 + y) + atanh(x / 2)",
         np.array([math.erf(v) for v in x]) + np.array([math.erfc(v) for v in y]) + np.arcsinh(x) + np.arccosh(1 + y) + np.arctanh(x / 2)),
        ("min(x, y, 0.5) + max(x, y, 0.9, 1.0) + sum(x, y, 1) + avg(x, y, 3) + min(x) + avg(x, y)",
         np.minimum(np.minimum(x, y), 0.5) + np.maximum(np.maximum(x, y), 1.0) + (x + y + 1) + (x + y + 3) / 3 + x + (x + y) / 2),
        ("if(x < 0.5 & y >= 0.4, 2, if(x > 0.8 | y == 7, 3, 4))",
         np.where((x < 0.5) & (y >= 0.4), 2.0, np.where(x > 0.8, 3.0, 4.0))),
    ]
    for expr, want in cases:
        got, td = eval_expr(expr, pts)
        assert not td, expr
        np.testing.assert_allclose(got, want, rtol=4e-16, atol=1e-300, err_msg=expr)

    with pytest.raises(Exception, match="hexadecimal"):
        eval_expr("0x10 + x", pts)
    # deal.II's random built-ins: rand_seed(s) is a reproducible uniform [0,1) stream per seed, rand() a clock-seeded one
    a, _ = eval_expr("rand_seed(7) + 0*x", pts)
    b, _ = eval_expr("rand_seed(8) + 0*x", pts)
    r, _ = eval_expr("rand() + 0*x", pts)
    for v in (a, b, r):
        assert v.min() >= 0.0 and v.max() < 1.0 and len(np.unique(v)) == len(v)
    assert not np.array_equal(a, b)
    with pytest.raises(Exception, match="no argument"):
        eval_expr("rand(1)", pts)
    got, _ = eval_expr("k * sin(2*pi*x) + Pi + big_name_2", pts, constants="pi=3.1415926535, k = 0.6, big_name_2=-1")
    np.testing.assert_allclose(got, 0.6 * np.sin(2 * math.pi * x) + math.pi - 1, rtol=4e-16)
    got, td = eval_expr("x*t + y", pts, t=0.25)
    assert td
    np.testing.assert_allclose(got, x * 0.25 + y, rtol=4e-16)
    got, _ = eval_expr("x + 10*y + 100*z", rng.random((5, 3)), dim=3)
    assert got.shape == (5,)
    got, td = eval_expr("x*t", np.array([[2.0]]), t=3.0, dim=1)
    assert td and got[0] == 6.0


@pytest.mark.parametrize("text, message", [
    ("set foo = 1", "no entry with name 'foo'"),
    ("set n_dims = 4", "outside the allowed range"),
    ("set n_dims = two", "is not an integer"),
    ("set fe_degree = 7", "outside the allowed range"),               # Patterns::Integer(1, 6), five_moment.h:116
    ("set t_end = -1", "outside the allowed range"),                  # Patterns::Double(0.0)
    ("set write_output = maybe", "is not a boolean"),
    ("set fields_enabled = sometimes", "is not one of true|false|auto"),
    ("set Application = Vlasov", "is not one of FiveMoment|FPETest"),
    ("subsection geometry\n set GridType = Extension\nend", "needs a grid extension"),
    ("subsection geometry\n set GridType = ForwardFacingStep\nend", "not supported by the GPU path"),
    ("subsection geometry\n set nx = 3", "unbalanced"),
    ("end", "'end' without 'subsection'"),
    ("subsection Nowhere\nend", "no subsection 'Nowhere/'"),
    ("set n_dims = 2\nsubsection geometry\n set nx = 4\nend", "need 2 entries"),
    ("subsection geometry\n set periodic_dimensions = x,w\nend", "not one of x|y|z"),
    ("subsection Species_2\n set name = ion\nend", "no subsection 'Species_2/'"),     # n_species defaults to 1
    ("subsection Species_1\n subsection BoundaryConditions\n set 0 = Wall\n end\nend", "no entry with name"),  # n_boundaries = 0
    ("set n_boundaries = 1\nsubsection Species_1\n subsection BoundaryConditions\n set 0 = Sticky\n end\nend", "Wall|Outflow|Inflow"),
    ("subsection Species_1\n set name = photon\nend", "neutral|ion|electron"),
    ("subsection Species_1\n subsection InitialCondition\n set Function expression = 1; 2; 3\n end\nend", "number of expressions (3)"),
    ("subsection Species_1\n subsection InitialCondition\n set Function expression = 1; 2; 3; 4; q\n end\nend", "unknown identifier 'q'"),
    ("subsection Species_1\n subsection InitialCondition\n set Function expression = 1; 2; 3; 4; sin(x\n end\nend", "expected ')'"),
    ("subsection Species_1\n subsection InitialCondition\n set Function expression = 1; 2; 3; 4; frob(x)\n end\nend", "unknown function 'frob'"),
    ("subsection Species_1\n subsection InitialCondition\n set Function expression = 1; 2; 3; 4; x +\n end\nend", "unexpected end"),
    ("subsection Species_1\n subsection InitialCondition\n set Function constants = k\n end\nend", "expected name=value"),
    ("subsection Species_1\n subsection InitialCondition\n set VariablesType = Entropy\n end\nend", "Primitive|Conserved"),
])
def test_malformed_inputs_fail_with_a_message(text, message):
    with pytest.raises(WarpiiGpuError) as err:
        App(text)
    assert message in str(err.value), str(err.value)


def test_comments_defaults_and_workdir():
    a = App("# only a comment\n\n   set t_end = 1.5   # trailing comment\n")
    assert (a.n_dims, a.fe_degree, a.t_end, a.write_output, a.nx, a.left, a.right) == (1, 2, 1.5, True, [1], [0.0], [1.0])
    assert a.format_workdir("sod_shocktube") == "FiveMoment__sod_shocktube"           # WorkDir default %A__%I, warpii.cc:139-149
    assert App("set WorkDir = runs/%I-%A-%I").format_workdir("STDIN") == "runs/STDIN-FiveMoment-STDIN"
    # UtilitiesTests.RemoveFileExtensionTest (test/utilities_test.cc:4-8): directories and the last extension go
    for name, stem in [("foo.inp", "foo"), ("baz.foo.inp", "baz.foo"), ("../examples/foo.inp", "foo"), ("plain", "plain")]:
        assert a.format_workdir(name) == "FiveMoment__" + stem
    # later `set` lines and repeated subsections override earlier ones
    b = App("set fe_degree = 3\nsubsection geometry\n set nx = 5\nend\nset fe_degree = 4\nsubsection geometry\n set nx = 9\nend")
    assert (b.fe_degree, b.nx) == (4, [9])


def test_cli_reports_usage_and_input_errors(tmp_path):
    import subprocess
    exe = os.path.join(os.path.dirname(INPUTS), "..", "..", "warpii_b200", "bin", "warpii_gpu")
    exe = os.path.normpath(exe)
    if not os.path.exists(exe):
        pytest.skip("warpii_gpu not built")
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "Usage:" in r.stdout and "--setup-only" in r.stdout
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "no input source was requested" in r.stdout          # warpii.cc:53-57
    r = subprocess.run([exe, str(tmp_path / "missing.inp")], capture_output=True, text=True)
    assert r.returncode == 1 and "Could not open requested input file" in r.stderr    # warpii.cc:109-114
    bad = tmp_path / "bad.inp"
    bad.write_text("set n_dims = 9\n")
    r = subprocess.run([exe, str(bad)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 1 and "outside the allowed range" in r.stderr


@pytest.mark.parametrize("dim,p", [(1, 2), (2, 3), (3, 2)])
def test_vtu_frames(tmp_path, dim, p):
    """The frame writer (stand-in for output_results + DataOut, five_moment.h:245-315): names, derived fields, sub-cells."""
    import oracle
    from oracle import Oracle
    import dgsem_cases as cases
    from warpii_b200.capi import write_vtu
    nx = [3, 2, 2][:dim]
    o = Oracle(dim, p, nx, [0.0] * dim, [1.0, 2.0, 0.5][:dim], gamma=1.4, n_species=2, fields_enabled=True)
    rng = np.random.default_rng(dim)
    u = np.zeros(o.shape)
    for s in range(2):
        prim = np.concatenate([rng.uniform(0.5, 2.0, o.shape[:1] + (o.NN, 1)), rng.normal(size=o.shape[:1] + (o.NN, 3)),
                               rng.uniform(0.5, 2.0, o.shape[:1] + (o.NN, 1))], axis=-1)
        cases.to_state(prim, 1.4, nc=18, species=s, u=u)
    u[:, 10:] = rng.normal(size=(o.n_elems, 8, o.NN))
    path = tmp_path / "solution_000.vtu"
    write_vtu(path, dim, p, u, o.node_coords(), ["ion", "electron"], fields_enabled=True, gas_gamma=1.4, owner_rank=3)
    v = cases.read_vtu(path)
    nsub, verts = p ** dim, 2 ** dim
    assert v["n_points"] == o.n_elems * o.NN and v["n_cells"] == o.n_elems * nsub
    assert (v["types"] == {1: 3, 2: 9, 3: 12}[dim]).all() and (v["owner"] == 3).all() and len(v["owner"]) == v["n_cells"]
    pts = v["Points"]
    assert np.array_equal(pts[:, :dim], o.node_coords().reshape(-1, dim)) and (pts[:, dim:] == 0).all()
    for s, name in enumerate(["ion", "electron"]):
        for k, comp in enumerate(["density", "x_momentum", "y_momentum", "z_momentum", "energy"]):
            assert np.array_equal(v[f"{name}_{comp}"], u[:, 5 * s + k].reshape(-1))
        q = u[:, 5 * s:5 * s + 5].transpose(0, 2, 1).reshape(-1, 5)
        pr = np.array([oracle.pressure(x, 1.4) for x in q])
        np.testing.assert_allclose(v[f"{name}_pressure"], pr, rtol=1e-14)
        np.testing.assert_allclose(v[f"{name}_y_velocity"], q[:, 2] / q[:, 0], rtol=1e-15)
        np.testing.assert_allclose(v[f"{name}_specific_entropy"], np.log(pr) - 1.4 * np.log(q[:, 0]), rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(v[f"{name}_speed_of_sound"], np.sqrt(1.4 * pr / q[:, 0]), rtol=1e-14)
    for k, name in enumerate(["E_field_x", "E_field_y", "E_field_z", "B_field_x", "B_field_y", "B_field_z",
                              "ph_maxwell_gauss_error", "ph_maxwell_monopole_error"]):
        assert np.array_equal(v[name], u[:, 10 + k].reshape(-1))
    cells = v["connectivity"].reshape(-1, verts)
    assert (v["offsets"] == verts * np.arange(1, len(cells) + 1)).all()
    assert (cells // o.NN == cells[:, :1] // o.NN).all()                      # a sub-cell never spans two elements
    # the sub-cells tile the box: their measures (from the first vertex and the far corner) add up to its volume
    lo, hi = pts[cells].min(axis=1)[:, :dim], pts[cells].max(axis=1)[:, :dim]
    assert np.isclose(np.prod(hi - lo, axis=1).sum(), np.prod([1.0, 2.0, 0.5][:dim]), rtol=1e-13)
    if dim >= 2:                                                               # counter-clockwise bottom face
        a, b, c = pts[cells[:, 0]], pts[cells[:, 1]], pts[cells[:, 2]]
        assert (((b - a)[:, 0] * (c - a)[:, 1] - (c - a)[:, 0] * (b - a)[:, 1]) > 0).all()
    if dim == 3:
        assert (pts[cells[:, 4], 2] > pts[cells[:, 0], 2]).all()


def test_gpus_launcher_does_not_hang_when_rank_0_dies_early(tmp_path):
    """`warpii_gpu --gpus N`: rank 0 hands the NCCL id to the launcher over a pipe, the launcher relays it.  If rank 0 dies
    before it has an id (here: no CUDA device in the CPU container, so its context cannot be created) every reader must see
    end-of-file and the launcher must come back with a non-zero exit code.  Each forked rank therefore closes every pipe
    end it inherited and does not use (warpii_cli.hpp); a rank that kept the write end of the upward pipe open made the
    launcher and all other ranks wait for ever."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU (rank 0 must fail before the id exists)")
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "warpii_b200", "bin", "warpii_gpu")
    (tmp_path / "run.inp").write_text(read_input("inflow_channel_2d.inp"))
    r = subprocess.run([exe, "--gpus", "3", "run.inp"], capture_output=True, text=True, cwd=tmp_path, timeout=60)
    assert r.returncode != 0
    assert "CUDA" in r.stderr or "cuda" in r.stderr, r.stderr[-500:]
