"""ctypes binding of tests/adapter/libadapter_check.so (TEST INFRASTRUCTURE): the adapter header include/gpu_es_dgsem_operator.h
compiled against the deal.II stand-in with the reference's own rk.h / solution_vec / bc_helper / dof_utils, driven the way
FiveMomentDGSolver::solve drives the operator."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "adapter", "_build", "libadapter_check.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available():
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "adapter")], stdout=subprocess.DEVNULL)
    return os.path.exists(_LIB)


def lib():
    L = C.CDLL(_LIB)
    L.adapter_last_error.restype = C.c_char_p
    L.adapter_run.argtypes = [C.c_int, C.c_int, _ip, _dp, _dp, _ip, C.c_int, C.c_int, C.c_double, _ip, _dp, _dp, C.c_int, _dp, _dp]
    return L


def run(dim, fe_degree, nx, left, right, periodic, state_cell_node_comp, n_steps, gamma, n_species=1, fields=False, bc_kind=None,
        inflow=None):
    """state_cell_node_comp: [cell lexicographic][node][comp] (the stub DoFHandler's numbering); returns (rc, error, state, t, bif)."""
    L = lib()
    ia = lambda v: (C.c_int * len(v))(*[int(x) for x in v])
    da = lambda v: (C.c_double * len(v))(*[float(x) for x in v])
    st = np.ascontiguousarray(state_cell_node_comp, dtype=np.float64).copy()
    bc = np.zeros(n_species * 2 * dim, dtype=np.int32) if bc_kind is None else np.ascontiguousarray(bc_kind, dtype=np.int32).reshape(-1)
    infl = np.zeros(n_species * 2 * dim * 5) if inflow is None else np.ascontiguousarray(inflow, dtype=np.float64).reshape(-1)
    t = C.c_double(0)
    bif = np.zeros(5 * 2 * dim)
    rc = L.adapter_run(dim, fe_degree, ia(nx), da(left), da(right), ia(periodic), n_species, int(fields), gamma,
                       bc.ctypes.data_as(_ip), infl.ctypes.data_as(_dp), st.ctypes.data_as(_dp), n_steps, C.byref(t), bif.ctypes.data_as(_dp))
    return rc, L.adapter_last_error().decode(), st, t.value, bif
