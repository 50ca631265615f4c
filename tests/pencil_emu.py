"""ctypes binding of tests/emu/libpencil_emu.so: the pencil stage kernel's phase functions executed on the host, thread by
thread and phase by phase (TEST INFRASTRUCTURE; see tests/emu/pencil_emu.cc).  Used by the CPU-only tests to check the
kernel source against the oracle; the product never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "emu", "_build", "libpencil_emu.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "emu")], stdout=subprocess.DEVNULL)
        L = C.CDLL(_LIB)
        L.emu_pencil_stage.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, _dp, _dp, _ip, _dp, _dp,
                                       _dp, _dp, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _dp, _dp, _dp, C.c_int,
                                       C.c_double, C.c_double, _dp, C.c_int, C.c_double, C.c_double, C.c_double]
        _lib = L
    return _lib


def patch_elems(dim, np_):
    return lib().emu_pencil_patch_elems(dim, np_)


def _p(a):
    return a.ctypes.data_as(_dp) if a is not None else C.cast(None, _dp)


def stage(dim, p, u, nbr, h, gamma, mode=1, dt=0.0, a=1.0, beta=0.0, dst=None, ghost=None, bres=None, elem_range=None, nsp=1,
          sol_in=None, dst2=None, sources=None, want_alpha=False, want_vmax=False, maxwell=None, field_kernel=False, fields_skip=False):
    """One launch.  u, dst: [n_elems][nc][NN] (dst is updated in place and returned); nbr: int32 [n_elems][2 dim].
    field_kernel=True runs the stand-alone field kernel (dgsem_maxwell_thread.cuh: the 8 field components only) instead of the
    pencil stage kernel."""
    L = lib()
    n_elems, nc, NN = u.shape
    assert NN == (p + 1) ** dim
    u = np.ascontiguousarray(u, dtype=np.float64)
    dst = np.zeros_like(u) if dst is None else dst
    nbr = np.ascontiguousarray(nbr, dtype=np.int32)
    b, e = (0, n_elems) if elem_range is None else elem_range
    alpha = np.zeros((n_elems, nsp)) if want_alpha else None
    vmax = np.zeros(1) if want_vmax else None
    hh = np.ascontiguousarray(list(h) + [1.0] * (3 - len(h)), dtype=np.float64)
    src_on, eps0, chi, qm = 0, 1.0, 0.0, None
    if sources:
        src_on, eps0, chi = 1, sources["epsilon0"], sources["chi"]
        qm = np.ascontiguousarray(sources["charge_over_mass"], dtype=np.float64)
    L.emu_select_field_kernel(1 if field_kernel else 0)
    rc = L.emu_pencil_stage(dim, p + 1, n_elems, b, e, nc, nsp, _p(u), _p(dst), nbr.ctypes.data_as(_ip), _p(ghost), _p(bres), _p(alpha),
                            _p(vmax), mode, gamma, dt, a, beta, _p(hh), _p(sol_in), _p(dst2), src_on, eps0, chi, _p(qm),
                            2 if fields_skip else (1 if maxwell else 0), (maxwell or {}).get("light_speed", 1.0), (maxwell or {}).get("chi", 0.0),
                            (maxwell or {}).get("gamma", 0.0))
    L.emu_select_field_kernel(0)
    assert rc == 0
    out = [dst]
    if want_alpha:
        out.append(alpha)
    if want_vmax:
        out.append(float(vmax[0]))
    return out[0] if len(out) == 1 else tuple(out)
