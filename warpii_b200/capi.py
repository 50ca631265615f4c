"""ctypes binding of include/warpii_gpu.h and include/warpii_host.h (no compute happens in Python)."""
import ctypes as C
import os

import numpy as np

BC_WALL, BC_OUTFLOW, BC_INFLOW, BC_SUBSONIC_OUTFLOW = 0, 1, 2, 3
INFLOW_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double), C.c_void_p)
FRAME_FN = C.CFUNCTYPE(None, C.c_uint, C.c_double, C.c_void_p)
FUSE_CFL = 1
NCCL_ID_BYTES = 128

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_i32p = C.POINTER(C.c_int32)
_i64p = C.POINTER(C.c_int64)


class WarpiiGpuError(RuntimeError):
    pass


def lib_path():
    # WARPII_B200_LIB selects another BUILD of the same library (tuning experiments); never a different backend
    return os.environ.get("WARPII_B200_LIB") or os.path.join(_HERE, "lib", "libwarpii_b200.so")


_lib = None


def lib():
    """Load libwarpii_b200.so; there is deliberately no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise WarpiiGpuError(
            f"{path} is missing: build it with `make -C warpii_b200` (or __graft_entry__.build()). "
            "warpii_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.warpii_gpu_last_error.restype = C.c_char_p
    L.warpii_host_last_error.restype = C.c_char_p
    L.warpii_gpu_abi_version.restype = C.c_int
    L.warpii_gpu_n_dofs.restype = C.c_int64
    L.warpii_gpu_n_dofs.argtypes = [vp]
    L.warpii_gpu_synchronize.argtypes = [vp]
    L.warpii_gpu_upload_state.argtypes = [vp, C.c_int, _dp, _i64p]
    L.warpii_gpu_download_state.argtypes = [vp, C.c_int, _dp, _i64p]
    L.warpii_gpu_zero_state.argtypes = [vp, C.c_int]
    L.warpii_gpu_copy_state.argtypes = [vp, C.c_int, C.c_int]
    L.warpii_gpu_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.warpii_gpu_set_inflow.argtypes = [vp, C.c_int, C.c_int, _dp]
    L.warpii_gpu_forward_euler_step.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    L.warpii_gpu_forward_euler_step_ex.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]
    L.warpii_gpu_recommend_dt.argtypes = [vp, C.c_int, _dp]
    L.warpii_gpu_max_transport_speed.argtypes = [vp, C.c_int, _dp]
    L.warpii_gpu_ssprk2_step.argtypes = [vp, C.c_int, C.c_int, C.c_double, C.c_double]
    L.warpii_gpu_advance_to.argtypes = [vp, C.c_int, C.c_int, _dp, C.c_double, C.c_double, C.c_int64, _i64p]
    L.warpii_gpu_boundary_fluxes.argtypes = [vp, C.c_int, _dp]
    L.warpii_gpu_set_boundary_fluxes.argtypes = [vp, C.c_int, _dp]
    L.warpii_gpu_global_integral.argtypes = [vp, C.c_int, C.c_int, _dp]
    L.warpii_gpu_shock_indicator.argtypes = [vp, C.c_int, _dp]
    L.warpii_gpu_rhs.argtypes = [vp, C.c_int, C.c_int, C.c_double]
    L.warpii_gpu_nccl_unique_id.argtypes = [C.c_char_p]
    L.warpii_gpu_launch_count.restype = C.c_int64
    L.warpii_gpu_launch_count.argtypes = [vp]
    L.warpii_gpu_stage_timing.argtypes = [vp, C.c_int, _dp, _i64p]
    L.warpii_gpu_stream.argtypes = [vp, C.POINTER(vp)]
    L.warpii_gpu_destroy.argtypes = [vp]
    L.warpii_box_solver_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _i32p, _dp, _dp, _i32p, C.c_int,
                                           _i32p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.warpii_box_solver_destroy.argtypes = [vp]
    L.warpii_box_solver_ctx.restype = vp
    L.warpii_box_solver_ctx.argtypes = [vp]
    for name in ("n_local_elems", "n_interface_elems", "n_ghost_faces"):
        f = getattr(L, "warpii_box_solver_" + name)
        f.restype = C.c_int64
        f.argtypes = [vp]
    L.warpii_box_solver_n_components.argtypes = [vp]
    L.warpii_box_solver_nodes_per_elem.argtypes = [vp]
    L.warpii_box_solver_local_to_global.argtypes = [vp, _i64p]
    L.warpii_box_solver_node_coords.argtypes = [vp, _dp]
    L.warpii_box_solver_set_state.argtypes = [vp, _dp]
    L.warpii_box_solver_get_state.argtypes = [vp, _dp]
    L.warpii_box_solver_set_inflow.argtypes = [vp, C.c_int, C.c_int, _dp]
    L.warpii_box_solver_attach_comm.argtypes = [vp, C.c_char_p]
    L.warpii_box_solver_set_inflow_function.argtypes = [vp, C.c_int, C.c_int, INFLOW_FN, vp, C.c_int]
    L.warpii_box_solver_set_sources.argtypes = [vp, C.c_int, C.c_double, C.c_double, _dp]
    L.warpii_box_solver_set_maxwell.argtypes = [vp, C.c_int, C.c_double, C.c_double, C.c_double]
    L.warpii_box_solver_n_boundary_faces.restype = C.c_int64
    L.warpii_box_solver_n_boundary_faces.argtypes = [vp]
    L.warpii_box_solver_boundary_points.argtypes = [vp, _dp, _i32p]
    L.warpii_gpu_n_boundary_points.argtypes = [vp, _i64p, _i32p]
    L.warpii_gpu_set_inflow_table.argtypes = [vp, C.c_int, _dp]
    L.warpii_app_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.warpii_app_destroy.argtypes = [vp]
    L.warpii_app_solver.restype = vp
    L.warpii_app_solver.argtypes = [vp]
    L.warpii_app_describe.argtypes = [vp, _i32p, _dp]
    L.warpii_app_species.argtypes = [vp, C.c_int, C.c_char_p, _dp, _dp, _i32p]
    L.warpii_app_eval_function.argtypes = [vp, C.c_int, C.c_int, C.c_int64, _dp, C.c_double, _dp, _i32p]
    L.warpii_app_set_output_dir.argtypes = [vp, C.c_char_p]
    L.warpii_app_format_workdir.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_int]
    L.warpii_app_set_device_loop.argtypes = [vp, C.c_int]
    L.warpii_app_setup.argtypes = [vp]
    L.warpii_app_run.argtypes = [vp, FRAME_FN, vp, _i64p]
    L.warpii_box_solver_solve.argtypes = [vp, C.c_double, C.c_double, C.c_double, vp, vp, _i64p]
    L.warpii_box_solver_step.argtypes = [vp, C.c_double, C.c_double]
    L.warpii_box_solver_recommend_dt.argtypes = [vp, _dp]
    L.warpii_box_solver_lsrk_step.argtypes = [vp, C.c_int, C.c_double, C.c_double, _dp, C.POINTER(C.c_int)]
    _lib = L
    return L


def _check(status, host=False):
    if status != 0:
        L = lib()
        msg = (L.warpii_host_last_error() if host else L.warpii_gpu_last_error()) or b""
        if host and not msg:
            msg = L.warpii_gpu_last_error() or b""
        raise WarpiiGpuError(msg.decode("utf-8", "replace"))


def _ptr(a):
    return a.ctypes.data_as(_dp)


def _arr32(v):
    return np.ascontiguousarray(v, dtype=np.int32)


def nccl_unique_id():
    buf = C.create_string_buffer(NCCL_ID_BYTES)
    _check(lib().warpii_gpu_nccl_unique_id(buf))
    return buf.raw


STEP_FN = C.CFUNCTYPE(C.c_int, C.c_double, C.c_double, C.c_void_p)
DT_FN = C.CFUNCTYPE(C.c_double, C.c_void_p)
CBI_FN = C.CFUNCTYPE(None, C.c_double, C.c_int, C.c_void_p)
CB_FN = C.CFUNCTYPE(None, C.c_double, C.c_void_p)


def host_advance(step, t_end, recommend_dt, callbacks):
    """The product's advance() (warpii_b200/host/gpu_operator.hpp); callbacks = [(interval, fn, zeroth, final)]."""
    L = lib()
    L.warpii_host_advance.argtypes = [STEP_FN, C.c_double, DT_FN, C.c_int, _dp, _i32p, _i32p, CBI_FN, C.c_void_p]
    n = len(callbacks)
    iv = np.array([c[0] for c in callbacks] or [0.0], dtype=np.float64)
    pz = _arr32([int(c[2]) for c in callbacks] or [0])
    pf = _arr32([int(c[3]) for c in callbacks] or [0])
    s = STEP_FN(lambda t, dt, _u: 1 if step(t, dt) else 0)
    d = DT_FN(lambda _u: recommend_dt())
    cb = CBI_FN(lambda t, i, _u: callbacks[i][1](t))
    _check(L.warpii_host_advance(s, t_end, d, n, _ptr(iv), pz.ctypes.data_as(_i32p), pf.ctypes.data_as(_i32p), cb, None), host=True)


def lsrk_coefficients(scheme):
    """(b, a, c) of the host layer's LowStorageRungeKuttaIntegrator (no GPU needed)."""
    L = lib()
    L.warpii_box_solver_lsrk_step.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, _dp, C.POINTER(C.c_int)]
    buf = np.zeros(32)
    n = C.c_int(0)
    _check(L.warpii_box_solver_lsrk_step(None, scheme, 0.0, 0.0, _ptr(buf), C.byref(n)), host=True)
    n = n.value
    return buf[:n].copy(), buf[n:2 * n - 1].copy(), buf[2 * n - 1:3 * n - 1].copy()


def elems_per_block(dim, fe_degree):
    return lib().warpii_gpu_elems_per_block(dim, fe_degree)


def box_tables(dim, nx, periodic, rank=0, n_ranks=1, group=1):
    """Mesh tables of one rank (no GPU needed): dict of numpy arrays.  group = patch size of the numbering."""
    L = lib()
    L.warpii_host_box_tables.argtypes = [C.c_int, _i32p, _i32p, C.c_int, C.c_int, C.c_int, _i64p, _i64p, _i32p, _i32p, _i32p,
                                         _i32p, _i32p, _i64p, _i64p, _i32p, _i32p, _i64p, _i32p]
    nx_a, per_a = _arr32(nx), _arr32([int(bool(p)) for p in periodic])
    counts = np.zeros(6, dtype=np.int64)
    none32, none64 = C.cast(None, _i32p), C.cast(None, _i64p)
    _check(L.warpii_host_box_tables(dim, nx_a.ctypes.data_as(_i32p), per_a.ctypes.data_as(_i32p), rank, n_ranks, group,
                                    counts.ctypes.data_as(_i64p), none64, none32, none32, none32, none32, none32, none64,
                                    none64, none32, none32, none64, none32), host=True)
    n_local, n_iface, n_ghost, n_bf, n_peers, n_send = [int(v) for v in counts]
    t = {
        "n_local": n_local, "n_interface": n_iface, "n_ghost": n_ghost, "n_bfaces": n_bf,
        "local_to_global": np.zeros(n_local, dtype=np.int64),
        "face_neighbor": np.zeros((n_local, 2 * dim), dtype=np.int32),
        "bf_elem": np.zeros(n_bf, dtype=np.int32), "bf_side": np.zeros(n_bf, dtype=np.int32),
        "bf_id": np.zeros(n_bf, dtype=np.int32), "peer_rank": np.zeros(n_peers, dtype=np.int32),
        "send_offset": np.zeros(n_peers + 1, dtype=np.int64), "recv_offset": np.zeros(n_peers + 1, dtype=np.int64),
        "send_elem": np.zeros(n_send, dtype=np.int32), "send_side": np.zeros(n_send, dtype=np.int32),
        "ghost_global_elem": np.zeros(n_ghost, dtype=np.int64), "ghost_side": np.zeros(n_ghost, dtype=np.int32),
    }
    p32 = lambda k: t[k].ctypes.data_as(_i32p)
    p64 = lambda k: t[k].ctypes.data_as(_i64p)
    _check(L.warpii_host_box_tables(dim, nx_a.ctypes.data_as(_i32p), per_a.ctypes.data_as(_i32p), rank, n_ranks, group,
                                    counts.ctypes.data_as(_i64p), p64("local_to_global"), p32("face_neighbor"), p32("bf_elem"),
                                    p32("bf_side"), p32("bf_id"), p32("peer_rank"), p64("send_offset"), p64("recv_offset"),
                                    p32("send_elem"), p32("send_side"), p64("ghost_global_elem"), p32("ghost_side")), host=True)
    return t


class BoxSolver:
    """FiveMomentGpuSolver on a box grid (warpii_b200/host/dg_solver.hpp) through the C ABI.

    Vector 0 is the solution, vector 1 the SSPRK2 scratch f_1.  State arrays are [local elem][comp][node].
    """

    def __init__(self, dim, fe_degree, nx, left, right, periodic=None, gamma=1.6666666666667, n_species=1,
                 fields_enabled=False, n_boundaries=None, bc_kinds=None, rank=0, n_ranks=1, device=0):
        L = lib()
        periodic = [1] * dim if periodic is None else [int(bool(p)) for p in periodic]
        if n_boundaries is None:
            n_boundaries = 0 if all(periodic) else 2 * dim
        self.dim, self.p, self.gamma, self.nsp, self.n_boundaries = dim, fe_degree, gamma, n_species, n_boundaries
        nx_a, per_a = _arr32(nx), _arr32(periodic)
        l_a = np.ascontiguousarray(left, dtype=np.float64)
        r_a = np.ascontiguousarray(right, dtype=np.float64)
        bc_p = C.cast(None, _i32p)
        if bc_kinds is not None and n_boundaries > 0:
            self._bc = _arr32(np.asarray(bc_kinds).reshape(n_species, n_boundaries))
            bc_p = self._bc.ctypes.data_as(_i32p)
        h = C.c_void_p()
        _check(L.warpii_box_solver_create(dim, fe_degree, n_species, int(fields_enabled), gamma, nx_a.ctypes.data_as(_i32p),
                                          _ptr(l_a), _ptr(r_a), per_a.ctypes.data_as(_i32p), n_boundaries, bc_p, rank, n_ranks,
                                          device, C.byref(h)), host=True)
        self._owned = True
        self._bind(h)

    @classmethod
    def mapped(cls, dim, fe_degree, nx, left, right, mapping, periodic=None, gamma=1.6666666666667, n_species=1,
               fields_enabled=False, n_boundaries=None, bc_kinds=None, rank=0, n_ranks=1, device=0):
        """The solver on a mapped box (curved elements, general-geometry kernels): mapping(x: ndarray[dim]) -> ndarray[dim]
        is applied to every Gauss-Lobatto support point (warpii_mapped_box_solver_create)."""
        L = lib()
        MAP_FN = C.CFUNCTYPE(None, _dp, _dp, C.c_void_p)
        L.warpii_mapped_box_solver_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, _i32p, _dp, _dp, _i32p, C.c_int,
                                                      _i32p, MAP_FN, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        periodic = [1] * dim if periodic is None else [int(bool(p)) for p in periodic]
        if n_boundaries is None:
            n_boundaries = 0 if all(periodic) else 2 * dim

        def thunk(x, y, _user):
            out = mapping(np.array([x[d] for d in range(dim)]))
            for d in range(dim):
                y[d] = float(out[d])

        self = cls.__new__(cls)
        self.dim, self.p, self.gamma, self.nsp, self.n_boundaries = dim, fe_degree, gamma, n_species, n_boundaries
        nx_a, per_a = _arr32(nx), _arr32(periodic)
        l_a, r_a = np.ascontiguousarray(left, dtype=np.float64), np.ascontiguousarray(right, dtype=np.float64)
        bc_p = C.cast(None, _i32p)
        if bc_kinds is not None and n_boundaries > 0:
            self._bc = _arr32(np.asarray(bc_kinds).reshape(n_species, n_boundaries))
            bc_p = self._bc.ctypes.data_as(_i32p)
        cb = MAP_FN(thunk)
        h = C.c_void_p()
        _check(L.warpii_mapped_box_solver_create(dim, fe_degree, n_species, int(fields_enabled), gamma, nx_a.ctypes.data_as(_i32p),
                                                 _ptr(l_a), _ptr(r_a), per_a.ctypes.data_as(_i32p), n_boundaries, bc_p, cb, None,
                                                 rank, n_ranks, device, C.byref(h)), host=True)
        self._owned = True
        self._bind(h)
        return self

    @classmethod
    def _view(cls, h, dim, fe_degree, gamma, n_species, n_boundaries):
        """A BoxSolver over a handle owned by something else (App.solver)."""
        self = cls.__new__(cls)
        self.dim, self.p, self.gamma, self.nsp, self.n_boundaries = dim, fe_degree, gamma, n_species, n_boundaries
        self._owned = False
        self._bind(h)
        return self

    def _bind(self, h):
        L = lib()
        self.h = h
        self.ctx = C.c_void_p(L.warpii_box_solver_ctx(h))
        self.n_elems = L.warpii_box_solver_n_local_elems(h)
        self.nc = L.warpii_box_solver_n_components(h)
        self.NN = L.warpii_box_solver_nodes_per_elem(h)
        self.shape = (self.n_elems, self.nc, self.NN)
        self.n_dofs = self.n_elems * self.nc * self.NN
        self.l2g = self.local_to_global()   # device order -> global lexicographic element index

    # ---- global (lexicographic) element order <-> this rank's device order -----------------------------
    def upload_global(self, vec, u_global):
        self.upload(vec, np.ascontiguousarray(u_global[self.l2g]))

    def download_global(self, vec, out=None):
        """Scatter this rank's elements into a global-order array (other ranks' elements are left untouched)."""
        u = self.download(vec)
        if out is None:
            assert self.n_elems == int(self.l2g.max()) + 1, "pass `out` on a sharded run"
            out = np.empty_like(u)
        out[self.l2g] = u
        return out

    def set_state_global(self, u_global):
        self.set_state(np.ascontiguousarray(u_global[self.l2g]))

    def get_state_global(self, out=None):
        u = self.get_state()
        if out is None:
            out = np.empty_like(u)
        out[self.l2g] = u
        return out

    def shock_indicator_global(self, vec=0):
        a = self.shock_indicator(vec)
        out = np.empty_like(a)
        out[self.l2g] = a
        return out

    def close(self):
        if getattr(self, "h", None):
            if self._owned:
                lib().warpii_box_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    # ---- host layer ---------------------------------------------------------------------------------
    def local_to_global(self):
        out = np.zeros(self.n_elems, dtype=np.int64)
        _check(lib().warpii_box_solver_local_to_global(self.h, out.ctypes.data_as(_i64p)), host=True)
        return out

    def node_coords(self):
        xyz = np.zeros((self.n_elems, self.NN, self.dim))
        _check(lib().warpii_box_solver_node_coords(self.h, _ptr(xyz)), host=True)
        return xyz

    def set_state(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.size == self.n_dofs
        _check(lib().warpii_box_solver_set_state(self.h, _ptr(u)), host=True)

    def get_state(self):
        u = np.zeros(self.shape)
        _check(lib().warpii_box_solver_get_state(self.h, _ptr(u)), host=True)
        return u

    def set_inflow(self, species, boundary_id, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        _check(lib().warpii_box_solver_set_inflow(self.h, species, boundary_id, _ptr(q)), host=True)

    def set_inflow_function(self, species, boundary_id, fn, time_dependent=True):
        """fn(x: ndarray[dim], t) -> 5 conserved values (EulerBCMap::set_inflow_boundary, bc_helper.h:52-64)."""
        dim = self.dim

        def thunk(x, t, q5, _user):
            q = fn(np.array([x[d] for d in range(dim)]), t)
            for k in range(5):
                q5[k] = float(q[k])

        cb = INFLOW_FN(thunk)
        if not hasattr(self, "_inflow_cbs"):
            self._inflow_cbs = []
        self._inflow_cbs.append(cb)   # keep the thunk alive as long as the solver
        _check(lib().warpii_box_solver_set_inflow_function(self.h, species, boundary_id, cb, None, int(time_dependent)), host=True)

    def set_maxwell(self, enabled, light_speed=1.0, chi=0.0, gamma=0.0):
        """Perfectly hyperbolic Maxwell fluxes for the field components (not in the reference operator): warpii_gpu_set_maxwell."""
        _check(lib().warpii_box_solver_set_maxwell(self.h, int(enabled), light_speed, chi, gamma), host=True)

    def set_sources(self, enabled, epsilon0=1.0, chi=0.0, charge_over_mass=None):
        """Two-fluid source terms (north_star kernel 4; not in the reference operator): warpii_gpu_set_sources."""
        qm = np.ascontiguousarray(charge_over_mass if charge_over_mass is not None else np.zeros(self.nsp), dtype=np.float64)
        _check(lib().warpii_box_solver_set_sources(self.h, int(enabled), epsilon0, chi, _ptr(qm)), host=True)

    def boundary_points(self):
        """(xyz[face][point][dim], boundary id per face) of this rank's boundary quadrature points."""
        nf = lib().warpii_box_solver_n_boundary_faces(self.h)
        nq = (self.p + 2) ** (self.dim - 1)
        xyz = np.zeros((nf, nq, self.dim))
        ids = np.zeros(nf, dtype=np.int32)
        _check(lib().warpii_box_solver_boundary_points(self.h, _ptr(xyz), ids.ctypes.data_as(_i32p)), host=True)
        return xyz, ids

    def set_inflow_table(self, species, table):
        table = np.ascontiguousarray(table, dtype=np.float64)
        _check(lib().warpii_gpu_set_inflow_table(self.ctx, species, _ptr(table)))

    def attach_comm(self, nccl_id):
        _check(lib().warpii_box_solver_attach_comm(self.h, nccl_id), host=True)

    def solve(self, t_end, fixed_dt=0.0, callback=None, callback_interval=0.0):
        steps = C.c_int64(0)
        cb = CB_FN(lambda t, _u: callback(t)) if callback else None
        _check(lib().warpii_box_solver_solve(self.h, t_end, fixed_dt, callback_interval, C.cast(cb, C.c_void_p) if cb else None,
                                             None, C.byref(steps)), host=True)
        return steps.value

    def global_error(self, exact, component=0, species=0):
        """compute_global_error (dg_solution_helper.cc:50-69): exact(x: ndarray[dim]) -> 5 values."""
        dim = self.dim

        def thunk(x, _t, q5, _user):
            v = exact(np.array([x[d] for d in range(dim)]))
            for k in range(5):
                q5[k] = float(v[k])

        cb = INFLOW_FN(thunk)
        L = lib()
        L.warpii_box_solver_global_error.argtypes = [C.c_void_p, INFLOW_FN, C.c_void_p, C.c_int, C.c_int, _dp]
        out = C.c_double(0)
        _check(L.warpii_box_solver_global_error(self.h, cb, None, species, component, C.byref(out)), host=True)
        return out.value

    def lsrk_step(self, scheme, dt, t=0.0):
        """LowStorageRungeKuttaIntegrator(scheme).perform_time_step on the solver's solution (host layer); the solution may
        live in another device vector afterwards: read it with get_state()."""
        _check(lib().warpii_box_solver_lsrk_step(self.h, scheme, dt, t, None, None), host=True)

    # ---- operator ABI ---------------------------------------------------------------------------------
    def upload(self, vec, u):
        u = np.ascontiguousarray(u, dtype=np.float64)
        assert u.size == self.n_dofs
        _check(lib().warpii_gpu_upload_state(self.ctx, vec, _ptr(u), None))

    def download(self, vec):
        u = np.zeros(self.shape)
        _check(lib().warpii_gpu_download_state(self.ctx, vec, _ptr(u), None))
        return u

    def zero(self, vec):
        _check(lib().warpii_gpu_zero_state(self.ctx, vec))

    def forward_euler_step(self, dst, u, dt, t, alpha=1.0, beta=0.0, flags=0):
        _check(lib().warpii_gpu_forward_euler_step_ex(self.ctx, dst, u, dt, t, alpha, beta, flags))

    def rhs(self, dst, u, t=0.0):
        _check(lib().warpii_gpu_rhs(self.ctx, dst, u, t))

    def ssprk2_step(self, dt, t, solution=0, f1=1):
        _check(lib().warpii_gpu_ssprk2_step(self.ctx, solution, f1, dt, t))

    def advance_to(self, t, t_stop, fixed_dt=0.0, max_steps=0, solution=0, f1=1):
        tt = C.c_double(t)
        steps = C.c_int64(0)
        _check(lib().warpii_gpu_advance_to(self.ctx, solution, f1, C.byref(tt), t_stop, fixed_dt, max_steps, C.byref(steps)))
        return tt.value, steps.value

    def lsrk_stage(self, sol_out, r_out, sol_in, r_in, factor_solution, factor_ai, t=0.0):
        L = lib()
        L.warpii_gpu_lsrk_stage.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]
        _check(L.warpii_gpu_lsrk_stage(self.ctx, sol_out, r_out, sol_in, r_in, factor_solution, factor_ai, t))

    def host_step(self, host_in, host_out, dt, t=0.0, solution=0, f1=1, n_slabs=0):
        """warpii_gpu_host_ssprk2_step: one SSPRK2 step of a state in (pinned) host memory, transfers overlapped with the
        stages; returns recommend_dt of the new state."""
        L = lib()
        L.warpii_gpu_host_ssprk2_step.argtypes = [C.c_void_p, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_double, _dp, C.c_int]
        assert host_in.dtype == np.float64 and host_out.dtype == np.float64 and host_in.size == self.n_dofs == host_out.size
        nxt = C.c_double(0)
        _check(L.warpii_gpu_host_ssprk2_step(self.ctx, solution, f1, _ptr(host_in), _ptr(host_out), dt, t, C.byref(nxt), n_slabs))
        return nxt.value

    def recommend_dt(self, vec=0):
        dt = C.c_double(0)
        _check(lib().warpii_gpu_recommend_dt(self.ctx, vec, C.byref(dt)))
        return dt.value

    def max_transport_speed(self, vec=0):
        v = C.c_double(0)
        _check(lib().warpii_gpu_max_transport_speed(self.ctx, vec, C.byref(v)))
        return v.value

    def boundary_fluxes(self, vec=0):
        out = np.zeros(5 * max(self.n_boundaries, 1))
        _check(lib().warpii_gpu_boundary_fluxes(self.ctx, vec, _ptr(out)))
        return out[:5 * self.n_boundaries]

    def global_integral(self, vec=0, species=0):
        out = np.zeros(5)
        _check(lib().warpii_gpu_global_integral(self.ctx, vec, species, _ptr(out)))
        return out

    def shock_indicator(self, vec=0):
        a = np.zeros((self.n_elems, self.nsp))
        _check(lib().warpii_gpu_shock_indicator(self.ctx, vec, _ptr(a)))
        return a

    def synchronize(self):
        _check(lib().warpii_gpu_synchronize(self.ctx))

    def launch_count(self):
        return lib().warpii_gpu_launch_count(self.ctx)

    def stage_timing(self, enable=True):
        ms = C.c_double(0)
        n = C.c_int64(0)
        _check(lib().warpii_gpu_stage_timing(self.ctx, int(enable), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def sm_clock_probes(self):
        """SM clock readings (MHz) taken on the device after every batch of steps while stage timing was enabled."""
        out = np.zeros(64)
        n = C.c_int(0)
        L = lib()
        L.warpii_gpu_sm_clock_probes.argtypes = [C.c_void_p, _dp, C.c_int, C.POINTER(C.c_int)]
        _check(L.warpii_gpu_sm_clock_probes(self.ctx, _ptr(out), 64, C.byref(n)))
        return out[:n.value].tolist()

    def device_ptr(self, vec):
        p = C.c_void_p()
        _check(lib().warpii_gpu_device_ptr(self.ctx, vec, C.byref(p)))
        return p.value

    def stream(self):
        p = C.c_void_p()
        _check(lib().warpii_gpu_stream(self.ctx, C.byref(p)))
        return p.value


class _Mesh(C.Structure):   # warpii_gpu_mesh
    _fields_ = [("dim", C.c_int32), ("fe_degree", C.c_int32), ("n_species", C.c_int32), ("fields_enabled", C.c_int32),
                ("gas_gamma", C.c_double), ("n_elems", C.c_int64), ("n_ghost_faces", C.c_int64),
                ("n_boundary_faces", C.c_int64), ("n_boundaries", C.c_int32), ("h", C.c_double * 3),
                ("face_neighbor", _i32p), ("boundary_face_elem", _i32p), ("boundary_face_side", _i32p),
                ("boundary_face_id", _i32p), ("bc_kind", _i32p), ("n_vectors", C.c_int32)]


class _Geometry(C.Structure):   # warpii_gpu_geometry
    _fields_ = [("inverse_jacobian", _dp), ("face_normal", _dp), ("face_jacobian", _dp), ("neighbor_face", _i32p),
                ("boundary_normal", _dp), ("boundary_jacobian", _dp)]


def mapped_metrics(dim, fe_degree, xyz, face_neighbor, neighbor_face=None, bf_elem=None, bf_side=None):
    """warpii_host_mapped_metrics (warpii_b200/host/mapped_mesh.hpp): the tables of warpii_gpu_geometry from the elements'
    Gauss-Lobatto support points xyz[n_elems][Np^dim][dim].  No GPU needed."""
    L = lib()
    L.warpii_host_mapped_metrics.argtypes = [C.c_int, C.c_int, C.c_int64, _dp, _i32p, _i32p, C.c_int64, _i32p, _i32p, _dp, _dp,
                                             _dp, _dp, _dp, _dp]
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    n_elems, NN = xyz.shape[0], xyz.shape[1]
    Np = fe_degree + 1
    NF, NG = Np ** (dim - 1), (Np + 1) ** (dim - 1)
    nbr = _arr32(face_neighbor)
    nbf = None if neighbor_face is None else _arr32(neighbor_face)
    bfe = _arr32(bf_elem if bf_elem is not None else [])
    bfs = _arr32(bf_side if bf_side is not None else [])
    nb = bfe.size
    g = {"inverse_jacobian": np.zeros((n_elems, NN, dim, dim)), "face_normal": np.zeros((n_elems, 2 * dim, NF, dim)),
         "face_jacobian": np.zeros((n_elems, 2 * dim, NF)), "boundary_normal": np.zeros((nb, NG, dim)),
         "boundary_jacobian": np.zeros((nb, NG)), "boundary_points": np.zeros((nb, NG, dim))}
    _check(L.warpii_host_mapped_metrics(dim, fe_degree, n_elems, _ptr(xyz), nbr.ctypes.data_as(_i32p),
                                        nbf.ctypes.data_as(_i32p) if nbf is not None else None, nb,
                                        bfe.ctypes.data_as(_i32p), bfs.ctypes.data_as(_i32p), _ptr(g["inverse_jacobian"]),
                                        _ptr(g["face_normal"]), _ptr(g["face_jacobian"]), _ptr(g["boundary_normal"]),
                                        _ptr(g["boundary_jacobian"]), _ptr(g["boundary_points"])), host=True)
    return g


class MeshSolver(BoxSolver):
    """The operator ABI (include/warpii_gpu.h) on caller-supplied mesh tables and general geometry: warpii_gpu_create +
    warpii_gpu_set_geometry.  Elements are in the caller's order; vector 0 = solution, 1 = f_1."""

    def __init__(self, dim, fe_degree, mesh, geometry, n_boundaries=0, bc_kinds=None, gamma=1.6666666666667, n_species=1,
                 fields_enabled=False, device=0, n_vectors=2):
        L = lib()
        L.warpii_gpu_create.argtypes = [C.POINTER(_Mesh), C.c_int, C.POINTER(C.c_void_p)]
        L.warpii_gpu_set_geometry.argtypes = [C.c_void_p, C.POINTER(_Geometry)]
        self.dim, self.p, self.gamma, self.nsp, self.n_boundaries = dim, fe_degree, gamma, n_species, n_boundaries
        keep = self._keep = {}
        keep["nbr"] = _arr32(mesh["face_neighbor"])
        keep["bfe"], keep["bfs"], keep["bfi"] = (_arr32(mesh.get(k, [])) for k in ("bf_elem", "bf_side", "bf_id"))
        keep["bc"] = _arr32(np.asarray(bc_kinds if bc_kinds is not None else np.zeros(n_species * max(n_boundaries, 1))).reshape(-1))
        m = _Mesh()
        m.dim, m.fe_degree, m.n_species, m.fields_enabled, m.gas_gamma = dim, fe_degree, n_species, int(fields_enabled), gamma
        m.n_elems, m.n_ghost_faces, m.n_boundary_faces, m.n_boundaries = keep["nbr"].shape[0], 0, keep["bfe"].size, n_boundaries
        m.h[0] = m.h[1] = m.h[2] = 1.0   # unused once the geometry is set
        p32 = lambda a: a.ctypes.data_as(_i32p)
        m.face_neighbor, m.boundary_face_elem, m.boundary_face_side = p32(keep["nbr"]), p32(keep["bfe"]), p32(keep["bfs"])
        m.boundary_face_id, m.bc_kind, m.n_vectors = p32(keep["bfi"]), p32(keep["bc"]), n_vectors
        ctx = C.c_void_p()
        _check(L.warpii_gpu_create(C.byref(m), device, C.byref(ctx)))
        self.ctx, self.h, self._owned = ctx, None, True
        g = _Geometry()
        for k in ("inverse_jacobian", "face_normal", "face_jacobian", "boundary_normal", "boundary_jacobian"):
            keep[k] = np.ascontiguousarray(geometry[k], dtype=np.float64)
            setattr(g, k, _ptr(keep[k]) if keep[k].size else None)
        nbf = mesh.get("neighbor_face")
        if nbf is not None:
            keep["nbf"] = _arr32(nbf)
            g.neighbor_face = p32(keep["nbf"])
        _check(L.warpii_gpu_set_geometry(self.ctx, C.byref(g)))
        self.n_elems = int(m.n_elems)
        self.nc = 5 * n_species + (8 if fields_enabled else 0)
        self.NN = (fe_degree + 1) ** dim
        self.shape = (self.n_elems, self.nc, self.NN)
        self.n_dofs = self.n_elems * self.nc * self.NN
        self.l2g = np.arange(self.n_elems, dtype=np.int64)

    def set_inflow(self, species, boundary_id, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        _check(lib().warpii_gpu_set_inflow(self.ctx, species, boundary_id, _ptr(q)))

    def set_sources(self, enabled, epsilon0=1.0, chi=0.0, charge_over_mass=None):
        L = lib()
        L.warpii_gpu_set_sources.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, _dp]
        qm = np.ascontiguousarray(charge_over_mass if charge_over_mass is not None else np.zeros(self.nsp), dtype=np.float64)
        _check(L.warpii_gpu_set_sources(self.ctx, int(enabled), epsilon0, chi, _ptr(qm)))

    def close(self):
        if getattr(self, "ctx", None):
            lib().warpii_gpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        self.close()


class App:
    """The FiveMoment application from a WarpII input file (include/warpii_host.h, warpii_app_*): the GPU-path stand-in
    for `Warpii::create_from_cli / setup / run` (warpii.cc:61-196) with Application = FiveMoment."""

    def __init__(self, input_text, rank=0, n_ranks=1, device=0):
        h = C.c_void_p()
        _check(lib().warpii_app_create(input_text.encode(), rank, n_ranks, device, C.byref(h)), host=True)
        self._adopt(h)

    def _adopt(self, h):
        self.h = h
        ints = np.zeros(16, dtype=np.int32)
        dbls = np.zeros(16)
        _check(lib().warpii_app_describe(h, ints.ctypes.data_as(_i32p), _ptr(dbls)), host=True)
        (self.n_dims, self.n_species, self.n_boundaries, self.fe_degree) = (int(v) for v in ints[:4])
        self.fields_enabled, self.write_output = bool(ints[4]), bool(ints[5])
        self.n_writeout_frames = int(ints[6])
        d = self.n_dims
        self.nx = [int(v) for v in ints[7:7 + d]]
        self.periodic = [bool(v) for v in ints[10:10 + d]]
        self.gas_gamma, self.t_end = float(dbls[0]), float(dbls[1])
        self.left = [float(v) for v in dbls[2:2 + d]]
        self.right = [float(v) for v in dbls[5:5 + d]]
        self.solver = None
        self.frames = []

    @classmethod
    def with_triangulation(cls, input_text, vertices, cells, face_boundary_ids=None, device=0):
        """GridType = Extension: the triangulation a GridExtension would populate, as arrays (warpii_host.h)."""
        L = lib()
        L.warpii_app_create_with_triangulation.argtypes = [C.c_char_p, C.c_int64, _dp, C.c_int64, _i32p, _i32p, C.c_int,
                                                           C.POINTER(C.c_void_p)]
        v = np.ascontiguousarray(vertices, dtype=np.float64)
        c = _arr32(cells)
        ids = None if face_boundary_ids is None else _arr32(face_boundary_ids)
        h = C.c_void_p()
        _check(L.warpii_app_create_with_triangulation(input_text.encode(), v.shape[0], _ptr(v), c.shape[0], c.ctypes.data_as(_i32p),
                                                      ids.ctypes.data_as(_i32p) if ids is not None else None, device, C.byref(h)),
               host=True)
        self = cls.__new__(cls)
        self._adopt(h)
        return self

    def species(self, i):
        name = C.create_string_buffer(16)
        charge, mass = C.c_double(), C.c_double()
        kinds = np.zeros(max(self.n_boundaries, 1), dtype=np.int32)
        _check(lib().warpii_app_species(self.h, i, name, C.byref(charge), C.byref(mass), kinds.ctypes.data_as(_i32p)), host=True)
        return dict(name=name.value.decode(), charge=charge.value, mass=mass.value, bc_kinds=[int(k) for k in kinds[:self.n_boundaries]])

    def eval_function(self, species, xyz, t=0.0, boundary_id=-1):
        """Conserved values of the parsed initial condition (boundary_id < 0) or inflow function at xyz[n][n_dims]."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1, self.n_dims)
        out = np.zeros((xyz.shape[0], 5))
        td = np.zeros(1, dtype=np.int32)
        _check(lib().warpii_app_eval_function(self.h, species, boundary_id, xyz.shape[0], _ptr(xyz), t, _ptr(out),
                                              td.ctypes.data_as(_i32p)), host=True)
        return out, bool(td[0])

    def format_workdir(self, input_name):
        buf = C.create_string_buffer(512)
        _check(lib().warpii_app_format_workdir(self.h, input_name.encode(), buf, 512), host=True)
        return buf.value.decode()

    def set_output_dir(self, path):
        _check(lib().warpii_app_set_output_dir(self.h, (path or "").encode()), host=True)

    def set_device_loop(self, on):
        _check(lib().warpii_app_set_device_loop(self.h, int(on)), host=True)

    def setup(self):
        _check(lib().warpii_app_setup(self.h), host=True)
        self.solver = BoxSolver._view(C.c_void_p(lib().warpii_app_solver(self.h)), self.n_dims, self.fe_degree, self.gas_gamma,
                                      self.n_species, self.n_boundaries)
        return self

    def run(self, frame_callback=None):
        """Returns the number of SSPRK2 steps; frame times end up in self.frames (frame 0 fires in setup)."""
        def on_frame(frame, t, _user):
            self.frames.append((int(frame), float(t)))
            if frame_callback:
                frame_callback(int(frame), float(t))

        cb = FRAME_FN(on_frame)
        steps = C.c_int64(0)
        _check(lib().warpii_app_run(self.h, cb, None, C.byref(steps)), host=True)
        if self.solver is None:
            self.solver = BoxSolver._view(C.c_void_p(lib().warpii_app_solver(self.h)), self.n_dims, self.fe_degree,
                                          self.gas_gamma, self.n_species, self.n_boundaries)
        return steps.value

    def close(self):
        if getattr(self, "h", None):
            if self.solver is not None:
                self.solver.close()
            lib().warpii_app_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()


def write_vtu(path, dim, fe_degree, state, xyz, species_names, fields_enabled=False, gas_gamma=5.0 / 3.0, owner_rank=0):
    """The product's frame writer on host arrays: state[elem][comp][node], xyz[elem][node][dim]."""
    L = lib()
    L.warpii_host_write_vtu.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_double,
                                        C.c_int, _dp, _dp]
    state = np.ascontiguousarray(state, dtype=np.float64)
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    _check(L.warpii_host_write_vtu(str(path).encode(), dim, fe_degree, state.shape[0], state.shape[1], len(species_names),
                                   ",".join(species_names).encode(), int(fields_enabled), gas_gamma, owner_rank, _ptr(state),
                                   _ptr(xyz)), host=True)


def check_division(n, seed=1, mode=0, device=0):
    """(mismatches, [a, b, got, want]) of div_rn_fast vs the IEEE division on n operand pairs (warpii_gpu_check_division)."""
    L = lib()
    L.warpii_gpu_check_division.argtypes = [C.c_int, C.c_int64, C.c_uint64, C.c_int, _i64p, _dp]
    count = C.c_int64(0)
    bad = np.zeros(4)
    _check(L.warpii_gpu_check_division(device, n, seed, mode, C.byref(count), _ptr(bad)))
    return count.value, bad


def point_fluxes(qa, qb, d, gamma, device=0):
    """Device evaluation of the EC flux (direction d) and the ES flux (normal +e_d) for state pairs; see warpii_gpu.h."""
    qa = np.ascontiguousarray(qa, dtype=np.float64).reshape(-1, 5)
    qb = np.ascontiguousarray(qb, dtype=np.float64).reshape(-1, 5)
    n = qa.shape[0]
    ec, es, prim = np.zeros((n, 5)), np.zeros((n, 5)), np.zeros((n, 12))
    L = lib()
    L.warpii_gpu_point_fluxes.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_double, _dp, _dp, _dp]
    _check(L.warpii_gpu_point_fluxes(device, n, _ptr(qa), _ptr(qb), d, gamma, _ptr(ec), _ptr(es), _ptr(prim)))
    return ec, es, prim
