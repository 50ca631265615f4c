"""The oracle against the REAL reference code for the rows that compile without deal.II.

oracle/_ref/libwarpii_ref.so is built by `make -C oracle ref` from the reference's own src/five_moment/euler.h,
src/tensor_utils.h, src/dof_utils.cc and src/timestepper.cc (where they lie under /root/reference) against the
minimal dealii::Tensor stand-in of oracle/ref_shim/.  The prebuilt library travels to the GPU box; if it is absent
(fresh clone without /root/reference) these tests skip.
"""
import ctypes as C
import os

import numpy as np
import pytest

import oracle

REF = os.path.join(os.path.dirname(os.path.abspath(oracle.__file__)), "_ref", "libwarpii_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference)")

_dp = C.POINTER(C.c_double)


def p(a):
    return np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(_dp)


@pytest.fixture(scope="module")
def ref():
    L = C.CDLL(REF)
    L.ref_ln_avg.restype = C.c_double
    L.ref_ln_avg.argtypes = [C.c_double, C.c_double]
    L.ref_pressure.restype = C.c_double
    L.ref_pressure.argtypes = [_dp, C.c_double]
    L.ref_euler_flux.argtypes = [C.c_int, _dp, C.c_double, _dp]
    L.ref_ec_flux.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp]
    L.ref_es_flux.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, _dp]
    L.ref_lf_flux.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, _dp]
    L.ref_entropy_variables.argtypes = [_dp, C.c_double, _dp]
    L.ref_mathematical_entropy.restype = C.c_double
    L.ref_mathematical_entropy.argtypes = [_dp, C.c_double]
    for n in ("ref_pencil_stride", "ref_pencil_base", "ref_quadrature_point_neighbor", "ref_quad_point_1d_index"):
        getattr(L, n).restype = C.c_uint
    # the reference takes ln from libm: compare with the oracle in the same mode
    oracle.set_log_impl(0)
    yield L
    oracle.set_log_impl(1)


def random_states(n, g):
    libc = C.CDLL("libc.so.6")
    libc.srand(1)
    r01 = lambda: (libc.rand() % 1000000000) / 1e9
    out = np.zeros((n, 5))
    for i in range(n):
        rho = (r01() + 1e-10) * 100
        u = [(r01() - 0.5) * 50 for _ in range(3)]
        pr = (r01() + 1e-10) * 100
        out[i] = [rho, rho * u[0], rho * u[1], rho * u[2], sum(0.5 * rho * v * v for v in u) + pr / (g - 1.0)]
    return out


def test_reference_goldens_on_the_reference_itself(ref):
    """test/euler_test.cc:70-116 evaluated by the compiled reference: the shim reproduces deal.II's Tensor semantics."""
    assert ref.ref_ln_avg(0.4, 0.4) == 0.4
    assert abs(ref.ref_ln_avg(1e-10, 1e-12) - 2.1497576854210972e-11) < 1e-16
    assert abs(ref.ref_ln_avg(1.0, 0.5) - 0.7213475204444817) < 1e-15
    g = 5.0 / 3.0
    left = [1.0, 0.0, 0.0, 0.0, 1.0 / (g - 1.0)]
    right = [0.1, 0.0, 0.0, 0.0, 0.125 / (g - 1.0)]
    out = np.zeros(5)
    ref.ref_es_flux(1, p(left), p(right), p([1.0]), g, p(out))
    assert abs(out[0] - 0.6495190528383291) < 1e-15 and abs(out[4] - 0.9381717944489488) < 1e-15
    F = np.zeros(5)
    ref.ref_ec_flux(1, p(left), p(right), g, p(F))
    assert F[0] == 0.0 and abs(F[1] - (0.5 + 1.0 / 9)) < 1e-15 and F[4] == 0.0


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_oracle_point_physics_is_bitwise_the_reference(ref, dim):
    """Same expressions in the same order: the restatement must reproduce the reference bit for bit."""
    g = 1.4
    s = random_states(300, g)
    rng = np.random.default_rng(3)
    for i in range(0, 300, 2):
        a, b = s[i], s[i + 1]
        n = rng.standard_normal(dim)
        n /= np.linalg.norm(n)
        assert ref.ref_ln_avg(a[0], b[0]) == oracle.ln_avg(a[0], b[0])
        assert ref.ref_pressure(p(a), g) == oracle.pressure(a, g)
        F = np.zeros((5, dim))
        ref.ref_euler_flux(dim, p(a), g, p(F))
        assert np.array_equal(F, oracle.euler_flux(dim, a, g))
        ref.ref_ec_flux(dim, p(a), p(b), g, p(F))
        assert np.array_equal(F, oracle.ec_flux(dim, a, b, g))
        out = np.zeros(5)
        ref.ref_es_flux(dim, p(a), p(b), p(n), g, p(out))
        assert np.array_equal(out, oracle.es_flux(dim, a, b, n, g))
        ref.ref_lf_flux(dim, p(a), p(b), p(n), g, p(out))
        assert np.array_equal(out, oracle.lf_flux(dim, a, b, n, g))
        w = np.zeros(5)
        ref.ref_entropy_variables(p(a), g, p(w))
        assert np.array_equal(w, oracle.entropy_variables(a, g))
        assert ref.ref_mathematical_entropy(p(a), g) == oracle.mathematical_entropy(a, g)


def test_oracle_near_equal_states_bitwise(ref):
    """The ill-conditioned regime of ln_avg (adjacent nodes of a fine mesh), still bit for bit."""
    g = 5.0 / 3.0
    rng = np.random.default_rng(5)
    base = random_states(50, g)
    for a in base:
        rho, u = a[0], a[1:4] / a[0]
        pr = (g - 1.0) * (a[4] - 0.5 * rho * (u @ u))
        for eps in (1e-3, 1e-5, 1e-6, 3e-7, 1e-9, 0.0):
            # perturb the primitive variables so that the pressure stays positive
            r2, u2, p2 = rho * (1 + eps * rng.standard_normal()), u * (1 + eps * rng.standard_normal(3)), pr * (1 + eps * rng.standard_normal())
            b = np.array([r2, *(r2 * u2), 0.5 * r2 * (u2 @ u2) + p2 / (g - 1.0)])
            F = np.zeros((5, 2))
            ref.ref_ec_flux(2, p(a), p(b), g, p(F))
            assert np.isfinite(F).all()
            assert np.array_equal(F, oracle.ec_flux(2, a, b, g))
    # a state with negative pressure poisons the flux with NaN in both (std::max semantics of ln_avg)
    bad = np.array([1.0, 3.0, 0.0, 0.0, 1.0])
    F = np.zeros((5, 2))
    ref.ref_ec_flux(2, p(base[0]), p(bad), g, p(F))
    assert np.array_equal(np.isnan(F), np.isnan(oracle.ec_flux(2, base[0], bad, g))) and np.isnan(F).any()


def test_oracle_index_maps_are_the_reference(ref):
    L = oracle.lib()
    for Np in range(2, 8):
        for d in range(2):
            assert ref.ref_pencil_stride(Np, d) == L.orc_pencil_stride(Np, d)
            buf_r, buf_o = (C.c_uint * 64)(), (C.c_uint * 64)()
            assert ref.ref_pencil_starts(2, Np, d, buf_r) == L.orc_pencil_starts(2, Np, d, buf_o)
            assert list(buf_r)[:Np] == list(buf_o)[:Np]
            for q in range(Np * Np):
                assert ref.ref_pencil_base(2, q, Np, d) == L.orc_pencil_base(2, q, Np, d)
                assert ref.ref_quad_point_1d_index(2, q, Np, d) == L.orc_quad_point_1d_index(2, q, Np, d)
                for k in range(Np):
                    assert ref.ref_quadrature_point_neighbor(2, q, k, Np, d) == L.orc_quadrature_point_neighbor(2, q, k, Np, d)
        for q in range(Np):
            assert ref.ref_pencil_base(1, q, Np, 0) == L.orc_pencil_base(1, q, Np, 0) == 0
            assert ref.ref_quad_point_1d_index(1, q, Np, 0) == L.orc_quad_point_1d_index(1, q, Np, 0)


def _run_advance(fn_lib, name, seq, t_end, cbs):
    STEP = C.CFUNCTYPE(C.c_int, C.c_double, C.c_double, C.c_void_p)
    DT = C.CFUNCTYPE(C.c_double, C.c_void_p)
    CB = C.CFUNCTYPE(None, C.c_double, C.c_int, C.c_void_p)
    log, i = [], [0]

    def dt(_u):
        i[0] += 1
        return seq[i[0] % len(seq)]

    f = getattr(fn_lib, name)
    f.argtypes = [STEP, C.c_double, DT, C.c_int, _dp, C.POINTER(C.c_int), C.POINTER(C.c_int), CB, C.c_void_p]
    n = len(cbs)
    iv = (C.c_double * max(n, 1))(*[c[0] for c in cbs])
    z = (C.c_int * max(n, 1))(*[int(c[1]) for c in cbs])
    fin = (C.c_int * max(n, 1))(*[int(c[2]) for c in cbs])
    f(STEP(lambda t, d, _u: (log.append(("step", t, d)) or 1)), t_end, DT(dt), n, iv, z, fin,
      CB(lambda t, k, _u: log.append(("cb", k, t))), None)
    return log


def test_time_loops_are_the_reference(ref):
    """Oracle advance(), and the PRODUCT's advance() (warpii_b200/host), against the compiled src/timestepper.cc."""
    import warpii_b200
    for seq, t_end, cbs in [([0.024], 1.2, [(0.3, True, True), (0.1, True, True), (0.25, False, True)]),
                            ([0.013, 0.0071, 0.02, 0.0033], 0.5, [(0.11, True, False)]),
                            ([0.05], 0.3, [])]:
        want = _run_advance(ref, "ref_advance", seq, t_end, cbs)
        assert len([e for e in want if e[0] == "step"]) > 0
        assert _run_advance(oracle.lib(), "orc_advance", seq, t_end, cbs) == want
        got = _run_advance(warpii_b200.lib(), "warpii_host_advance", seq, t_end, [(c[0], c[1], c[2]) for c in cbs])
        assert got == want


# ---- the reference's volume and subcell-FV DRIVERS ------------------------------------------------------
# SplitFormVolumeFlux<dim>::calculate_flux (split_form_volume_flux.h:61-99) and SubcellFiniteVolumeFlux<dim>::calculate_flux
# (subcell_finite_volume_flux.h:68-159), compiled from the reference sources against the one-cell FEEvaluation /
# VectorizedArray / FullMatrix stand-in of oracle/ref_shim/deal.II/matrix_free/fe_evaluation.h, evaluate ONE cell; the
# oracle's cell_residual must give the same integrated residual, on Cartesian and on curved cells.
def _cell_state(rng, NN, gamma, amp=0.2):
    prim = np.zeros((NN, 5))
    prim[:, 0] = 1.0 + amp * rng.uniform(-1, 1, NN)
    prim[:, 1:4] = 0.4 * rng.uniform(-1, 1, (NN, 3))
    prim[:, 4] = 1.0 + amp * rng.uniform(-1, 1, NN)
    cons = oracle.primitive_to_conserved(prim, gamma) if hasattr(oracle, "primitive_to_conserved") else None
    return np.ascontiguousarray(np.asarray(cons).T)     # [5][NN]


@pytest.mark.parametrize("dim,fe_degree,h", [(1, 2, [0.3]), (1, 4, [1.7]), (2, 1, [0.5, 0.25]), (2, 3, [0.02, 0.05]), (2, 4, [1.0, 1.0])])
@pytest.mark.parametrize("alpha", [0.0, 0.37])
def test_volume_and_fv_drivers_on_a_cartesian_cell(ref, dim, fe_degree, h, alpha):
    ref.ref_cell_residual.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp, C.c_double, C.c_int, _dp]
    ref.ref_diff_matrix.argtypes = [C.c_int, _dp]
    Np, gamma = fe_degree + 1, 1.4
    NN = Np ** dim
    D = np.zeros((Np, Np))
    ref.ref_diff_matrix(Np, p(D))
    assert np.abs(D - oracle.diff_matrix(Np)).max() <= 2e-14 * np.abs(D).max()       # D(j,l) = shape_grad(l, x_j)[0]
    o = oracle.Oracle(dim, fe_degree, [2] * dim, [0.0] * dim, [2 * v for v in h], gamma=gamma)
    rng = np.random.default_rng(11 * dim + fe_degree)
    ue = _cell_state(rng, NN, gamma)
    jinv = np.zeros((NN, dim, dim))
    for d in range(dim):
        jinv[:, d, d] = 1.0 / h[d]
    want = np.zeros((5, NN))
    assert ref.ref_cell_residual(dim, Np, gamma, p(ue), p(jinv), alpha, 3, p(want)) == 0
    got = o.cell_residual(ue, alpha)
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-14 * scale, np.abs(got - want).max() / scale
    # the two drivers separately: the subcell scheme is evaluated (and multiplied by alpha) whatever alpha is
    vol, fv = np.zeros((5, NN)), np.zeros((5, NN))
    ref.ref_cell_residual(dim, Np, gamma, p(ue), p(jinv), alpha, 1, p(vol))
    ref.ref_cell_residual(dim, Np, gamma, p(ue), p(jinv), alpha, 2, p(fv))
    assert np.abs(vol + fv - want).max() <= 1e-15 * scale
    if alpha == 0.0:
        assert np.abs(fv).max() == 0.0


@pytest.mark.parametrize("fe_degree", [2, 3])
@pytest.mark.parametrize("alpha", [0.0, 0.5])
def test_volume_and_fv_drivers_on_a_curved_cell(ref, fe_degree, alpha):
    """Curved 2-D cells (MappingQ(p) geometry): metric terms averaged between the nodes of a pair
    (split_form_volume_flux.h:82-84) and subcell normals accumulated with Q (subcell_finite_volume_flux.h:101-106)."""
    import mesh_cases as mc
    from oracle import GeneralOracle
    from warpii_b200.capi import mapped_metrics
    ref.ref_cell_residual.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp, C.c_double, C.c_int, _dp]
    dim, gamma = 2, 5.0 / 3.0
    Np = fe_degree + 1
    NN = Np ** dim
    left, right = [0.0, 0.0], [1.0, 1.3]
    mesh, xyz = mc.mapped_box(dim, fe_degree, [3, 2], left, right, [1, 1], mc.wavy(left, right, 0.06))
    geo = mapped_metrics(dim, fe_degree, xyz, mesh["face_neighbor"])
    o = GeneralOracle(dim, fe_degree, mesh, geo, gamma=gamma)
    rng = np.random.default_rng(5 + fe_degree)
    ue = _cell_state(rng, NN, gamma)
    jinv = np.ascontiguousarray(np.asarray(geo["inverse_jacobian"]).reshape(-1, NN, dim, dim)[0])
    assert np.abs(jinv[:, 0, 1]).max() > 1e-3          # the cell really is curved
    want = np.zeros((5, NN))
    assert ref.ref_cell_residual(dim, Np, gamma, p(ue), p(jinv), alpha, 3, p(want)) == 0
    got = o.cell_residual(ue, alpha)                    # element 0 of the mesh
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-14 * scale, np.abs(got - want).max() / scale
