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


@pytest.mark.parametrize("scheme,n,order", [(0, 3, 3), (1, 5, 4)])
def test_coefficients_satisfy_the_order_conditions(scheme, n, order):
    b, a, c = capi.lsrk_coefficients(scheme)
    assert len(b) == n and len(a) == n - 1 and len(c) == n
    A = tableau(a, b)
    assert np.abs(A.sum(axis=1) - c).max() <= 1e-15
    conds = [b.sum() - 1, b @ c - 1 / 2, b @ c ** 2 - 1 / 3, b @ (A @ c) - 1 / 6]
    if order >= 4:
        conds += [b @ c ** 3 - 1 / 4, (b * c) @ (A @ c) - 1 / 8, b @ (A @ c ** 2) - 1 / 12, b @ (A @ (A @ c)) - 1 / 24]
    assert np.abs(conds).max() <= 1e-14


@pytest.mark.parametrize("scheme", [2, 3])
def test_unavailable_schemes_report_not_implemented(scheme):
    with pytest.raises(capi.WarpiiGpuError, match="ExcNotImplemented"):
        capi.lsrk_coefficients(scheme)


@pytest.mark.parametrize("scheme,order", [(0, 3), (1, 4)])
def test_observed_order_in_time(scheme, order):
    b, a, c = capi.lsrk_coefficients(scheme)
    o = Oracle(1, 3, [8], [0.0], [1.0], gamma=1.4)
    u0 = o.project(cases.sine_wave(amp=0.2))
    T = 0.02

    def run(n):
        u = u0.copy()
        for k in range(n):
            o.lsrk_step(u, b, a, c, T / n, k * T / n)
        return u

    ref = run(64)
    e1 = np.abs(run(2) - ref).max()
    e2 = np.abs(run(4) - ref).max()
    assert e1 > 1e-12   # still resolvable above round-off
    assert np.log2(e1 / e2) > order - 0.4
