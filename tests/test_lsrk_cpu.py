"""Low-storage Runge-Kutta integrator (rk.h:10-77) on the CPU: the host layer's coefficient tables
and the oracle's stage.  The reference takes the coefficients from deal.II (absent here), so they are pinned by what
defines them: the order conditions of the 2-register scheme, and the observed order of accuracy in time."""
import numpy as np
import pytest

import dgsem_cases as cases
from oracle import Oracle
from warpii_b200 import capi


def tableau(a, b):
    s = len(b)
    A = np.zeros((s, s))
    for i in range(s):
        for j in range(i):
            A[i, j] = a[j] if j == i - 1 else b[j]
    return A


@pytest.mark.parametrize("scheme,n,order", [(0, 3, 3), (1, 5, 4), (2, 7, 4), (3, 9, 5)])
def test_coefficients_satisfy_the_order_conditions(scheme, n, order):
    b, a, c = capi.lsrk_coefficients(scheme)
    assert len(b) == n and len(a) == n - 1 and len(c) == n
    A = tableau(a, b)
    assert np.abs(A.sum(axis=1) - c).max() <= 1e-15
    Ac, Ac2 = A @ c, A @ c ** 2
    conds = [b.sum() - 1, b @ c - 1 / 2, b @ c ** 2 - 1 / 3, b @ Ac - 1 / 6]
    if order >= 4:
        conds += [b @ c ** 3 - 1 / 4, (b * c) @ Ac - 1 / 8, b @ Ac2 - 1 / 12, b @ (A @ Ac) - 1 / 24]
    if order >= 5:   # the nine rooted trees of order five
        conds += [b @ c ** 4 - 1 / 5, (b * c ** 2) @ Ac - 1 / 10, (b * c) @ Ac2 - 1 / 15, (b * c) @ (A @ Ac) - 1 / 30,
                  b @ (Ac * Ac) - 1 / 20, b @ (A @ c ** 3) - 1 / 20, b @ (A @ (c * Ac)) - 1 / 40, b @ (A @ Ac2) - 1 / 60,
                  b @ (A @ (A @ Ac)) - 1 / 120]
    assert np.abs(conds).max() <= 1e-14


def test_scheme_out_of_range_is_refused():
    with pytest.raises(capi.WarpiiGpuError):
        capi.lsrk_coefficients(4)


@pytest.mark.parametrize("scheme,order,T,n", [(0, 3, 0.02, 2), (1, 4, 0.02, 2), (2, 4, 0.04, 2), (3, 5, 0.06, 2)])
def test_observed_order_in_time(scheme, order, T, n):
    b, a, c = capi.lsrk_coefficients(scheme)
    o = Oracle(1, 3, [8], [0.0], [1.0], gamma=1.4)
    u0 = o.project(cases.sine_wave(amp=0.2))

    def run(n):
        u = u0.copy()
        for k in range(n):
            o.lsrk_step(u, b, a, c, T / n, k * T / n)
        return u

    ref = run(64)
    e1 = np.abs(run(n) - ref).max()
    e2 = np.abs(run(2 * n) - ref).max()
    assert e2 > 1e-13   # still resolvable above round-off
    assert np.log2(e1 / e2) > order - 0.4
