"""ctypes binding of the CPU oracle (oracle/dgsem_oracle.cc).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(warpii_b200) must never import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libwarpii_ref.so")

BC_WALL, BC_OUTFLOW, BC_INFLOW, BC_SUBSONIC_OUTFLOW = 0, 1, 2, 3
INFLOW_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.c_void_p)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_up = C.POINTER(C.c_uint)


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference exists)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("dgsem_oracle.cc", "det_log.h"))):
        subprocess.check_call(["make", "-C", _HERE, "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


def _ptr(a):
    return a.ctypes.data_as(_dp)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_lib = None
_variants = {}


def lib(variant=None):
    """The oracle library; variant="fma" is the same source compiled with multiply-add contraction (liboracle_fma.so), used
    only to measure how far two legitimate FP64 builds of the algorithm are apart."""
    global _lib
    if variant is not None:
        if variant not in _variants:
            path = os.path.join(_HERE, "_build", f"liboracle_{variant}.so")
            if not os.path.exists(path):
                subprocess.check_call(["make", "-C", _HERE, f"_build/liboracle_{variant}.so"], stdout=subprocess.DEVNULL)
            _variants[variant] = _bind(C.CDLL(path))
        return _variants[variant]
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = _bind(C.CDLL(_LIB_PATH))
    return _lib


def _bind(L):
    L.orc_ln_avg.restype = C.c_double
    L.orc_ln_avg.argtypes = [C.c_double, C.c_double]
    L.orc_det_log.restype = C.c_double
    L.orc_det_log.argtypes = [C.c_double]
    L.orc_set_log_impl.argtypes = [C.c_int]
    L.orc_pressure.restype = C.c_double
    L.orc_pressure.argtypes = [_dp, C.c_double]
    L.orc_euler_flux.argtypes = [C.c_int, _dp, C.c_double, _dp]
    L.orc_lf_flux.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, _dp]
    L.orc_ec_flux.argtypes = [C.c_int, _dp, _dp, C.c_double, _dp]
    L.orc_es_flux.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, _dp]
    L.orc_entropy_variables.argtypes = [_dp, C.c_double, _dp]
    L.orc_mathematical_entropy.restype = C.c_double
    L.orc_mathematical_entropy.argtypes = [_dp, C.c_double]
    L.orc_entropy_flux.argtypes = [C.c_int, _dp, C.c_double, _dp]
    L.orc_primitive_to_conserved.argtypes = [_dp, C.c_double, _dp]
    for name in ("orc_pencil_stride",):
        getattr(L, name).restype = C.c_uint
        getattr(L, name).argtypes = [C.c_uint, C.c_uint]
    L.orc_pencil_base.restype = C.c_uint
    L.orc_pencil_base.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint]
    L.orc_quadrature_point_neighbor.restype = C.c_uint
    L.orc_quadrature_point_neighbor.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_uint]
    L.orc_quad_point_1d_index.restype = C.c_uint
    L.orc_quad_point_1d_index.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint]
    L.orc_pencil_starts.restype = C.c_int
    L.orc_pencil_starts.argtypes = [C.c_int, C.c_uint, C.c_uint, _up]
    L.orc_gll.argtypes = [C.c_int, _dp, _dp]
    L.orc_gauss.argtypes = [C.c_int, _dp, _dp]
    L.orc_diff_matrix.argtypes = [C.c_int, _dp]
    L.orc_legendre_analysis_1d.argtypes = [C.c_int, _dp]
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _ip, _dp, _dp, _ip, _ip]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_threads.argtypes = [C.c_void_p, C.c_int]
    L.orc_n_elems.restype = C.c_int64
    L.orc_n_elems.argtypes = [C.c_void_p]
    L.orc_n_dofs.restype = C.c_int64
    L.orc_n_dofs.argtypes = [C.c_void_p]
    L.orc_n_components.argtypes = [C.c_void_p]
    L.orc_nodes_per_elem.argtypes = [C.c_void_p]
    L.orc_n_boundaries.argtypes = [C.c_void_p]
    L.orc_node_coords.argtypes = [C.c_void_p, _dp]
    L.orc_set_inflow.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp]
    L.orc_set_inflow_function.argtypes = [C.c_void_p, C.c_int, C.c_int, INFLOW_FN, C.c_void_p]
    L.orc_rhs.argtypes = [C.c_void_p, _dp, C.c_double, _dp, _dp]
    L.orc_alpha.argtypes = [C.c_void_p, _dp, _dp]
    L.orc_cell_residual.argtypes = [C.c_void_p, _dp, C.c_double, _dp]
    L.orc_shock_indicator.restype = C.c_double
    L.orc_shock_indicator.argtypes = [C.c_void_p, _dp]
    L.orc_forward_euler_step.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_double, C.c_double,
                                         C.c_double, _dp, _dp]
    L.orc_max_transport_speed.restype = C.c_double
    L.orc_max_transport_speed.argtypes = [C.c_void_p, _dp]
    L.orc_recommend_dt.restype = C.c_double
    L.orc_recommend_dt.argtypes = [C.c_void_p, _dp]
    L.orc_ssprk2_step.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_double, _dp, _dp]
    L.orc_solve.restype = C.c_int64
    L.orc_solve.argtypes = [C.c_void_p, _dp, C.c_double, _dp, C.c_int64, C.c_double]
    L.orc_global_integral.argtypes = [C.c_void_p, _dp, C.c_int, _dp]
    return L


STEP_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.c_double, C.c_void_p)
DT_FN = C.CFUNCTYPE(C.c_double, C.c_void_p)
CB_FN = C.CFUNCTYPE(None, C.c_double, C.c_int, C.c_void_p)


def advance(step, t_end, recommend_dt, callbacks):
    """callbacks: list of (interval, fn(t), perform_zeroth, perform_final). Mirrors timestepper.cc:6-56."""
    L = lib()
    L.orc_advance.argtypes = [STEP_FN, C.c_double, DT_FN, C.c_int, _dp, _ip, _ip, CB_FN, C.c_void_p]
    n = len(callbacks)
    iv = (C.c_double * max(n, 1))(*[c[0] for c in callbacks])
    pz = (C.c_int * max(n, 1))(*[int(c[2]) for c in callbacks])
    pf = (C.c_int * max(n, 1))(*[int(c[3]) for c in callbacks])
    s = STEP_FN(lambda t, dt, _u: 1 if step(t, dt) else 0)
    d = DT_FN(lambda _u: recommend_dt())
    cb = CB_FN(lambda t, i, _u: callbacks[i][1](t))
    L.orc_advance(s, t_end, d, n, iv, pz, pf, cb, None)


# ---- point physics helpers -------------------------------------------------
def set_log_impl(deterministic):
    """1: oracle/det_log.h inside ln_avg (default), 0: libm's log like the reference."""
    lib().orc_set_log_impl(int(bool(deterministic)))


def det_log(x):
    return lib().orc_det_log(float(x))


def ln_avg(a, b):
    return lib().orc_ln_avg(a, b)


def pressure(q, gamma):
    return lib().orc_pressure(_ptr(_f64(q)), gamma)


def euler_flux(dim, q, gamma):
    F = np.zeros((5, dim))
    lib().orc_euler_flux(dim, _ptr(_f64(q)), gamma, _ptr(F))
    return F


def ec_flux(dim, qj, ql, gamma):
    F = np.zeros((5, dim))
    lib().orc_ec_flux(dim, _ptr(_f64(qj)), _ptr(_f64(ql)), gamma, _ptr(F))
    return F


def es_flux(dim, qj, ql, n, gamma):
    out = np.zeros(5)
    lib().orc_es_flux(dim, _ptr(_f64(qj)), _ptr(_f64(ql)), _ptr(_f64(n)), gamma, _ptr(out))
    return out


def lf_flux(dim, qin, qout, n, gamma):
    out = np.zeros(5)
    lib().orc_lf_flux(dim, _ptr(_f64(qin)), _ptr(_f64(qout)), _ptr(_f64(n)), gamma, _ptr(out))
    return out


def entropy_variables(q, gamma):
    w = np.zeros(5)
    lib().orc_entropy_variables(_ptr(_f64(q)), gamma, _ptr(w))
    return w


def mathematical_entropy(q, gamma):
    return lib().orc_mathematical_entropy(_ptr(_f64(q)), gamma)


def entropy_flux(dim, q, gamma):
    out = np.zeros(dim)
    lib().orc_entropy_flux(dim, _ptr(_f64(q)), gamma, _ptr(out))
    return out


def primitive_to_conserved(prim, gamma):
    """prim[..., 5] = [rho, ux, uy, uz, p] -> conserved [..., 5] (species_func.cc:15-28), elementwise."""
    prim = _f64(prim)
    out = np.empty_like(prim)
    flat_in = prim.reshape(-1, 5)
    flat_out = out.reshape(-1, 5)
    L = lib()
    if flat_in.shape[0] > 4096:
        # vectorised restatement with the same operation order
        rho = flat_in[:, 0]
        flat_out[:, 0] = rho
        ke = np.zeros_like(rho)
        for d in range(3):
            flat_out[:, d + 1] = rho * flat_in[:, d + 1]
            ke = ke + 0.5 * rho * flat_in[:, d + 1] * flat_in[:, d + 1]
        flat_out[:, 4] = ke + flat_in[:, 4] / (gamma - 1)
        return out
    for i in range(flat_in.shape[0]):
        L.orc_primitive_to_conserved(_ptr(flat_in[i]), gamma, _ptr(flat_out[i]))
    return out


def gll(Np):
    x = np.zeros(Np)
    w = np.zeros(Np)
    lib().orc_gll(Np, _ptr(x), _ptr(w))
    return x, w


def gauss(n):
    x = np.zeros(n)
    w = np.zeros(n)
    lib().orc_gauss(n, _ptr(x), _ptr(w))
    return x, w


def diff_matrix(Np):
    D = np.zeros((Np, Np))
    lib().orc_diff_matrix(Np, _ptr(D))
    return D


def legendre_analysis_1d(Np):
    V = np.zeros((Np, Np))
    lib().orc_legendre_analysis_1d(Np, _ptr(V))
    return V


class Oracle:
    """Discretised ES-DGSEM operator on a Cartesian box, CPU restatement of the reference."""

    def __init__(self, dim, fe_degree, nx, left, right, periodic=None, gamma=1.6666666666667,
                 n_species=1, fields_enabled=False, bc_kinds=None, threads=1, variant=None):
        L = self._L = lib(variant)
        self.dim, self.p, self.gamma, self.nsp = dim, fe_degree, gamma, n_species
        self.nx, self.left, self.right = list(nx), list(left), list(right)
        periodic = [1] * dim if periodic is None else [int(bool(x)) for x in periodic]
        nx_a = (C.c_int * dim)(*[int(v) for v in nx])
        l_a = (C.c_double * dim)(*[float(v) for v in left])
        r_a = (C.c_double * dim)(*[float(v) for v in right])
        p_a = (C.c_int * dim)(*periodic)
        bc_a = None
        if bc_kinds is not None:
            flat = np.asarray(bc_kinds, dtype=np.int32).reshape(-1)
            assert flat.size == n_species * 2 * dim
            bc_a = (C.c_int * flat.size)(*[int(v) for v in flat])
        self.h = L.orc_create(dim, fe_degree, n_species, int(fields_enabled), gamma, nx_a, l_a, r_a, p_a, bc_a)
        if not self.h:
            raise ValueError("orc_create failed")
        self.n_elems = L.orc_n_elems(self.h)
        self.nc = L.orc_n_components(self.h)
        self.NN = L.orc_nodes_per_elem(self.h)
        self.n_dofs = L.orc_n_dofs(self.h)
        self.n_boundaries = L.orc_n_boundaries(self.h)
        self.shape = (self.n_elems, self.nc, self.NN)
        if threads != 1:
            L.orc_set_threads(self.h, threads)

    def __del__(self):
        if getattr(self, "h", None):
            self._L.orc_destroy(self.h)
            self.h = None

    def set_threads(self, n):
        self._L.orc_set_threads(self.h, n)

    def node_coords(self):
        xyz = np.zeros((self.n_elems, self.NN, self.dim))
        self._L.orc_node_coords(self.h, _ptr(xyz))
        return xyz

    def set_inflow(self, species, boundary_id, q):
        self._L.orc_set_inflow(self.h, species, boundary_id, _ptr(_f64(q)))

    def set_sources(self, enabled, epsilon0=1.0, chi=0.0, charge_over_mass=None):
        """Two-fluid source terms (new physics, not in the reference): see dgsem_oracle.cc::add_sources."""
        qm = _f64(charge_over_mass if charge_over_mass is not None else np.zeros(self.nsp))
        L = self._L
        L.orc_set_sources.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, _dp]
        L.orc_set_sources(self.h, int(enabled), epsilon0, chi, _ptr(qm))

    def set_maxwell(self, enabled, light_speed=1.0, chi=0.0, gamma=0.0):
        """Perfectly hyperbolic Maxwell fluxes for the 8 field components (new physics): dgsem_oracle.cc::add_maxwell."""
        L = self._L
        L.orc_set_maxwell.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_double]
        if L.orc_set_maxwell(self.h, int(enabled), light_speed, chi, gamma) != 0:
            raise ValueError("set_maxwell: needs the 8 field components, a Cartesian box and light_speed > 0")

    def set_inflow_function(self, species, boundary_id, fn):
        """fn(x: ndarray[dim], t: float) -> 5 conserved values, evaluated at every boundary quadrature point with the
        stage time, like the Function<dim> of EulerBCMap::get_inflow (fluid_flux_es_dgsem_operator.h:139-144, 381-384)."""
        dim = self.dim

        def thunk(x, t, q5, _user):
            q = fn(np.array([x[d] for d in range(dim)]), t)
            for k in range(5):
                q5[k] = float(q[k])

        cb = INFLOW_FN(thunk)
        if not hasattr(self, "_inflow_cbs"):
            self._inflow_cbs = {}
        self._inflow_cbs[(species, boundary_id)] = cb   # keep the thunk alive
        self._L.orc_set_inflow_function(self.h, species, boundary_id, cb, None)

    def project(self, prim_fn, species=0, u=None, conserved=False):
        """Nodal interpolation of an IC given as fn(xyz[...,dim]) -> [...,5] (dg_solution_helper.cc:24-48)."""
        if u is None:
            u = np.zeros(self.shape)
        xyz = self.node_coords()
        vals = _f64(prim_fn(xyz))
        cons = vals if conserved else primitive_to_conserved(vals, self.gamma)
        u[:, 5 * species:5 * species + 5, :] = np.transpose(cons, (0, 2, 1))
        return u

    def rhs(self, u, t=0.0):
        u = _f64(u)
        dudt = np.zeros(self.shape)
        bif = np.zeros(5 * self.n_boundaries)
        self._L.orc_rhs(self.h, _ptr(u), t, _ptr(dudt), _ptr(bif))
        return dudt, bif

    def alpha(self, u):
        u = _f64(u)
        a = np.zeros((self.n_elems, self.nsp))
        self._L.orc_alpha(self.h, _ptr(u), _ptr(a))
        return a

    def cell_residual(self, ue, alpha):
        ue = _f64(ue)
        R = np.zeros((5, self.NN))
        self._L.orc_cell_residual(self.h, _ptr(ue), alpha, _ptr(R))
        return R

    def shock_indicator(self, v):
        return self._L.orc_shock_indicator(self.h, _ptr(_f64(v)))

    def forward_euler_step(self, dst, u, dt, t, a=1.0, beta=0.0, bif_dst=None, bif_u=None):
        assert dst.flags.c_contiguous and dst.dtype == np.float64
        u = _f64(u)
        self._L.orc_forward_euler_step(self.h, _ptr(dst), _ptr(u), dt, t, a, beta,
                                     _ptr(bif_dst) if bif_dst is not None else None,
                                     _ptr(bif_u) if bif_u is not None else None)
        return dst

    def lsrk_stage(self, sol, r_out, r_in, factor_solution, factor_ai, t=0.0):
        """k = M^-1 R(r_in); r_out = sol + factor_ai k; sol += factor_solution k (rk.h:53-71); r_out may be r_in."""
        L = self._L
        L.orc_lsrk_stage.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double]
        assert sol.flags.c_contiguous and r_out.flags.c_contiguous and r_in.flags.c_contiguous
        L.orc_lsrk_stage(self.h, _ptr(sol), _ptr(r_out), _ptr(r_in), factor_solution, factor_ai, t)

    def lsrk_step(self, u, b, a, c, dt, t=0.0):
        """LowStorageRungeKuttaIntegrator::perform_time_step (rk.h:53-71) with coefficients (b, a, c)."""
        ri = np.zeros_like(u)
        self.lsrk_stage(u, ri, u.copy(), b[0] * dt, a[0] * dt, t)
        for s in range(1, len(b)):
            self.lsrk_stage(u, ri, ri.copy(), b[s] * dt, 0.0 if s == len(b) - 1 else a[s] * dt, t + c[s] * dt)
        return u

    def max_transport_speed(self, u):
        return self._L.orc_max_transport_speed(self.h, _ptr(_f64(u)))

    def recommend_dt(self, u):
        return self._L.orc_recommend_dt(self.h, _ptr(_f64(u)))

    def ssprk2_step(self, u, dt, t, bif=None):
        assert u.flags.c_contiguous and u.dtype == np.float64
        f1 = np.zeros_like(u)
        bif_f1 = np.zeros(5 * self.n_boundaries) if bif is not None else None
        self._L.orc_ssprk2_step(self.h, _ptr(u), _ptr(f1), dt, t,
                              _ptr(bif) if bif is not None else None,
                              _ptr(bif_f1) if bif is not None else None)
        return u

    def solve(self, u, t_end, bif=None, max_steps=0, fixed_dt=0.0):
        assert u.flags.c_contiguous and u.dtype == np.float64
        return self._L.orc_solve(self.h, _ptr(u), t_end, _ptr(bif) if bif is not None else None, max_steps, fixed_dt)

    def global_integral(self, u, species=0):
        out = np.zeros(5)
        self._L.orc_global_integral(self.h, _ptr(_f64(u)), species, _ptr(out))
        return out


class GeneralOracle(Oracle):
    """The same operator on an arbitrary conforming mesh of (curved) quadrilaterals / hexahedra: orc_create_general.

    mesh: dict with face_neighbor [n_elems][2*dim] (neighbour or -1 - boundary face number), optional neighbor_face,
    bf_id [n_bfaces]; geometry: dict with inverse_jacobian, face_normal, face_jacobian, boundary_normal,
    boundary_jacobian (the tables of include/warpii_gpu.h::warpii_gpu_geometry).  PARITY UNPINNED (see dgsem_oracle.h).
    """

    def __init__(self, dim, fe_degree, mesh, geometry, n_boundaries=0, bc_kinds=None, gamma=1.6666666666667, n_species=1,
                 fields_enabled=False, threads=1):
        L = self._L = lib()
        i64p, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
        L.orc_create_general.restype = C.c_void_p
        L.orc_create_general.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int64, C.c_int, i64p, i32p,
                                         C.c_int64, i32p, _ip, _dp, _dp, _dp, _dp, _dp]
        self.dim, self.p, self.gamma, self.nsp = dim, fe_degree, gamma, n_species
        nbr = np.ascontiguousarray(mesh["face_neighbor"], dtype=np.int64)
        n_elems = nbr.shape[0]
        nbrf = mesh.get("neighbor_face")
        nbrf = None if nbrf is None else np.ascontiguousarray(nbrf, dtype=np.int32)
        bf_id = np.ascontiguousarray(mesh.get("bf_id", np.zeros(0)), dtype=np.int32)
        bc_a = None
        if bc_kinds is not None and n_boundaries > 0:
            flat = np.asarray(bc_kinds, dtype=np.int32).reshape(-1)
            assert flat.size == n_species * n_boundaries
            bc_a = (C.c_int * flat.size)(*[int(v) for v in flat])
        g = {k: _f64(v) for k, v in geometry.items() if v is not None}
        opt = lambda k: _ptr(g[k]) if k in g and g[k].size else None
        self.h = L.orc_create_general(dim, fe_degree, n_species, int(fields_enabled), gamma, n_elems, n_boundaries,
                                      nbr.ctypes.data_as(i64p), nbrf.ctypes.data_as(i32p) if nbrf is not None else None,
                                      bf_id.size, bf_id.ctypes.data_as(i32p), bc_a, _ptr(g["inverse_jacobian"]),
                                      _ptr(g["face_normal"]), _ptr(g["face_jacobian"]), opt("boundary_normal"),
                                      opt("boundary_jacobian"))
        if not self.h:
            raise ValueError("orc_create_general failed (bad arguments or unknown boundary id)")
        self.n_elems = L.orc_n_elems(self.h)
        self.nc = L.orc_n_components(self.h)
        self.NN = L.orc_nodes_per_elem(self.h)
        self.n_dofs = L.orc_n_dofs(self.h)
        self.n_boundaries = L.orc_n_boundaries(self.h)
        self.shape = (self.n_elems, self.nc, self.NN)
        if threads != 1:
            L.orc_set_threads(self.h, threads)

    def node_coords(self):
        raise NotImplementedError("a general mesh has no box coordinates; the caller holds xyz")

    def set_inflow_function(self, *a, **k):
        raise NotImplementedError("inflow functions are a box-grid service of the oracle")

    def project_xyz(self, xyz, prim_fn, species=0, u=None, conserved=False):
        if u is None:
            u = np.zeros(self.shape)
        vals = _f64(prim_fn(xyz))
        cons = vals if conserved else primitive_to_conserved(vals, self.gamma)
        u[:, 5 * species:5 * species + 5, :] = np.transpose(cons, (0, 2, 1))
        return u
