"""Known-answer tests of the DEVICE point physics (the same device functions the stage kernel calls), against the
reference's golden vectors (test/euler_test.cc:70-116) and against the CPU oracle."""
import ctypes

import numpy as np
import pytest

import oracle
from warpii_b200.capi import point_fluxes

pytestmark = pytest.mark.gpu


def sod_states(g):
    return [1.0, 0.0, 0.0, 0.0, 1.0 / (g - 1.0)], [0.1, 0.0, 0.0, 0.0, 0.125 / (g - 1.0)]


# test/euler_test.cc:82-101
def test_device_ec_flux_goldens():
    g = 5.0 / 3.0
    left, right = sod_states(g)
    ec, _, _ = point_fluxes([left], [right], 0, g)
    assert ec[0, 0] == 0.0 and abs(ec[0, 1] - (0.5 + 1.0 / 9)) < 1e-15 and ec[0, 4] == 0.0
    left = [1.0, 1.0, 0.0, 0.0, 0.5 * 1.0 + 1.0 / (g - 1.0)]
    right = [0.5, 0.5, 0.0, 0.0, 0.5 * 0.5 + 0.5 / (g - 1.0)]
    ec, _, _ = point_fluxes([left], [right], 0, g)
    assert abs(ec[0, 0] - 0.7213475204444817) < 1e-15
    assert abs(ec[0, 1] - 1.4713475204444817) < 1e-15
    assert abs(ec[0, 4] - 2.192695040888963) < 1e-15    # the reference's bound; the device returns the golden literal bit for bit (0x1.18aa3b295c17ep+1)


# test/euler_test.cc:103-116
def test_device_es_flux_goldens():
    g = 5.0 / 3.0
    left, right = sod_states(g)
    _, es, _ = point_fluxes([left], [right], 0, g)
    assert abs(es[0, 0] - 0.6495190528383291) < 1e-15
    assert abs(es[0, 1] - (0.5 + 1.0 / 9)) < 1e-15
    assert abs(es[0, 4] - 0.9381717944489488) < 1e-15


# test/euler_test.cc:70-80 (ln_avg through rho_ln = F_ec[0] / u_avg for unit velocities)
def test_device_ln_avg_goldens():
    g = 1.4

    def rho_ln(a, b):
        # equal unit velocity, pressures chosen freely: F_ec[0] = ln_avg(rho_a, rho_b) * 1
        qa = [a, a, 0.0, 0.0, 0.5 * a + 1.0 / (g - 1.0)]
        qb = [b, b, 0.0, 0.0, 0.5 * b + 1.0 / (g - 1.0)]
        return point_fluxes([qa], [qb], 0, g)[0][0, 0]

    # the flux carries ln_avg times u_avg = (a/a + b/b)/2, itself good to ~2 ulp: allow 4 ulp on top of the reference's bounds
    ulp4 = 4 * 2.2e-16
    assert abs(rho_ln(0.4, 0.4) - 0.4) <= 0.4 * ulp4
    assert abs(rho_ln(1e-10, 1e-12) - 2.1497576854210972e-11) < 1e-16
    assert abs(rho_ln(0.4, 0.4 + 1e-8) - (0.8 + 1e-8) / 2.0) < 1e-16 + 0.4 * ulp4
    assert abs(rho_ln(1.0, 0.5) - 0.7213475204444817) < 1e-15


def random_states(n, g, seed=1):
    """test/euler_test_helpers.h: rho in (0,100], u in [-25,25]^3, p in (0,100] from glibc rand()."""
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(seed)
    r01 = lambda: (libc.rand() % 1000000000) / 1e9
    out = np.zeros((n, 5))
    for i in range(n):
        rho = (r01() + 1e-10) * 100
        u = [(r01() - 0.5) * 50 for _ in range(3)]
        p = (r01() + 1e-10) * 100
        out[i] = [rho, rho * u[0], rho * u[1], rho * u[2], sum(0.5 * rho * v * v for v in u) + p / (g - 1.0)]
    return out


@pytest.mark.parametrize("d", [0, 1, 2])
def test_device_fluxes_match_oracle_on_random_states(d):
    g = 1.4
    s = random_states(400, g)
    qa, qb = s[:200], s[200:]
    ec, es, prim = point_fluxes(qa, qb, d, g)
    n = np.zeros(3)
    n[d] = 1.0
    for i in range(200):
        want_ec = oracle.ec_flux(3, qa[i], qb[i], g)[:, d]
        want_es = oracle.es_flux(3, qa[i], qb[i], n, g)
        scale = np.abs(want_ec).max()
        assert np.abs(ec[i] - want_ec).max() <= 1e-14 * scale
        assert np.abs(es[i] - want_es).max() <= 1e-14 * max(scale, np.abs(want_es).max())
    # the ill-conditioned inputs of ln_avg are bit-identical to the CPU path's
    p = np.array([oracle.pressure(q, g) for q in qa])
    assert np.array_equal(prim[:, 8], p)
    assert np.array_equal(prim[:, 4], qa[:, 0] / (2.0 * p))
    assert np.array_equal(prim[:, 5], np.array([oracle.det_log(x) for x in qa[:, 0]]))
    assert np.array_equal(prim[:, 6], np.array([oracle.det_log(x) for x in prim[:, 4]]))


def test_device_es_flux_is_entropy_dissipative():
    """test/euler_test.cc:226-282 on the device flux (normal = e_x)."""
    g = 1.4
    s = random_states(200, g, seed=7)
    qa, qb = s[:100], s[100:]
    _, es, _ = point_fluxes(qa, qb, 0, g)
    for i in range(100):
        wL, wR = oracle.entropy_variables(qa[i], g), oracle.entropy_variables(qb[i], g)
        fL, fR = oracle.euler_flux(1, qa[i], g)[:, 0], oracle.euler_flux(1, qb[i], g)[:, 0]
        qL, qR = oracle.entropy_flux(1, qa[i], g)[0], oracle.entropy_flux(1, qb[i], g)[0]
        psiL, psiR = wL @ fL - qL, wR @ fR - qR
        assert (wR - wL) @ es[i] - (psiR - psiL) <= 1e-9 * (abs(qL) + abs(qR))


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_branch_free_division_is_the_ieee_division(mode):
    """div_rn_fast (csrc/det_log.cuh) == __ddiv_rn bit for bit: 2^27 operand pairs per distribution and seed."""
    from warpii_b200.capi import check_division
    for seed in (1, 20261017):
        bad, first = check_division(1 << 27, seed=seed, mode=mode)
        assert bad == 0, f"{bad} mismatches, e.g. a={first[0]!r} b={first[1]!r} got={first[2]!r} want={first[3]!r}"


def test_device_logarithm_is_the_oracle_logarithm_bit_for_bit():
    """200k densities over 12 decades: log rho and log beta of the device == oracle.det_log of the same doubles."""
    rng = np.random.default_rng(5)
    n = 200000
    g = 1.4
    rho = np.exp(rng.uniform(-14, 14, n))
    u = rng.normal(size=(n, 3))
    p = np.exp(rng.uniform(-10, 10, n))
    q = np.zeros((n, 5))
    q[:, 0] = rho
    q[:, 1:4] = rho[:, None] * u
    q[:, 4] = 0.5 * rho * (u ** 2).sum(axis=1) + p / (g - 1)
    _, _, prim = point_fluxes(q, q, 0, g)
    want_rho = np.array([oracle.det_log(x) for x in q[:, 0]])
    want_beta = np.array([oracle.det_log(x) for x in prim[:, 4]])
    assert np.array_equal(prim[:, 5], want_rho)
    assert np.array_equal(prim[:, 6], want_beta)
    pr = np.array([oracle.pressure(x, g) for x in q])
    assert np.array_equal(prim[:, 8], pr) and np.array_equal(prim[:, 4], q[:, 0] / (2.0 * pr))
    # and it is a logarithm: < 1 ulp from the correctly rounded value
    ref = np.log(q[:, 0].astype(np.longdouble)).astype(np.float64)
    assert (np.abs(want_rho - ref) <= np.spacing(np.abs(ref))).all()
