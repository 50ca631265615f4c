"""The pencil stage kernel's SOURCE (warpii_b200/csrc/dgsem_pencil_stage.cuh) executed on the host, phase by phase
(tests/emu), against the oracle.  This is the CPU tier's check of the kernel's indexing and arithmetic: pencil ownership in
the three directions, the swizzled shared-memory planes, in-patch / outside / ghost pencil ends, partial patches, the
troubled-cell correction, the stage-update modes, the fused transport speed.  The GPU tier (-m gpu) runs the same source as
compiled by nvcc through the C ABI; this file never touches a GPU and the product never loads the emulation.
Tolerances: north_star's 1e-12 relative L2 per RHS component (dgsem_cases.rhs_error_bound)."""
import numpy as np
import pytest

import dgsem_cases as cases
import oracle
import pencil_emu as emu
from oracle import Oracle
from warpii_b200 import box_tables


def face_node_to_node(dim, Np, d, side, t):
    idx = [0, 0, 0]
    for a in range(dim):
        if a == d:
            continue
        idx[a] = t % Np
        t //= Np
    idx[d] = Np - 1 if side else 0
    return idx[0] + Np * (idx[1] + Np * idx[2])


def setup(dim, p, nx, left, right, gamma, n_species=1, fields=False, rank=0, n_ranks=1):
    o = Oracle(dim, p, nx, left, right, gamma=gamma, n_species=n_species, fields_enabled=fields, threads=4)
    tab = box_tables(dim, nx, [1] * dim, rank=rank, n_ranks=n_ranks, group=emu.patch_elems(dim, p + 1))
    h = [(r - l) / n for l, r, n in zip(left, right, nx)]
    return o, tab, h


def check(got, want, u, o, h, tol=1e-12):
    scale = cases.summand_scale(u, o.gamma, o.dim, h, oracle.diff_matrix(o.p + 1))
    err, bound = cases.rhs_error_and_bound(got, want, scale, tol)
    assert np.isfinite(got).all()
    assert (err <= bound).all(), f"abs L2 error per component {err}, bound {bound}, plain {cases.rel_l2_per_component(got, want)}"


RHS_CASES = [
    (2, 3, [8, 8], [0.0, -5.0], [10.0, 5.0], cases.isentropic_vortex(), 1.4),
    (2, 3, [7, 9], [0.0, -5.0], [10.0, 5.0], cases.isentropic_vortex(), 1.4),          # partial patches
    (2, 3, [16, 12], [0.0, -5.0], [10.0, 5.0], cases.isentropic_vortex(), 1.4),        # 2 x 3 full 8x4 patches
    (2, 3, [9, 5], [0.0, -5.0], [10.0, 5.0], cases.isentropic_vortex(), 1.4),          # one full patch + three partial ones
    (2, 3, [3, 2], [0.0, -5.0], [10.0, 5.0], cases.isentropic_vortex(), 1.4),          # a single partial patch
    (2, 2, [9, 5], [0.0, 0.0], [1.0, 0.5], cases.sine_wave(vel=(1.0, 1.0, 0.0), wave=(1, 2, 0)), 5.0 / 3.0),
    (2, 4, [5, 6], [0.0, 0.0], [1.0, 1.0], cases.smooth_blob_3d(), 5.0 / 3.0),
    (2, 1, [6, 4], [0.0, 0.0], [1.0, 1.0], cases.smooth_blob_3d(), 1.4),
    (3, 3, [5, 4, 6], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], cases.smooth_blob_3d(), 5.0 / 3.0),
    (3, 3, [4, 4, 4], [0.0, -5.0, -5.0], [10.0, 5.0, 5.0], cases.isentropic_vortex(), 1.4),
    (3, 2, [6, 6, 3], [0.0, 0.0, 0.0], [1.0, 2.0, 1.0], cases.smooth_blob_3d(0.1), 1.4),
    (3, 4, [3, 3, 3], [0.0, -5.0, -5.0], [10.0, 5.0, 5.0], cases.isentropic_vortex(), 1.4),
    (3, 1, [4, 3, 2], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], cases.smooth_blob_3d(0.1), 5.0 / 3.0),
    (3, 5, [2, 2, 2], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], cases.smooth_blob_3d(0.1), 5.0 / 3.0),
]


@pytest.mark.parametrize("dim,p,nx,left,right,ic,gamma", RHS_CASES)
def test_rhs_periodic(dim, p, nx, left, right, ic, gamma):
    o, tab, h = setup(dim, p, nx, left, right, gamma)
    u = o.project(ic)
    want, _ = o.rhs(u)
    l2g = tab["local_to_global"]
    got, alpha = emu.stage(dim, p, u[l2g], tab["face_neighbor"], h, gamma, mode=1, want_alpha=True)
    check(got, want[l2g], u[l2g], o, h)
    assert np.allclose(alpha, o.alpha(u)[l2g], rtol=0, atol=1e-12)


def blast(dim, width=0.02):
    def fn(xyz):
        s = xyz[..., 0] - 0.45 + (0.3 * (xyz[..., 1] - 0.5) if dim > 1 else 0.0) + (0.2 * (xyz[..., 2] - 0.5) if dim > 2 else 0.0)
        w = 0.5 * (1 - np.tanh(s / width))
        out = np.zeros(xyz.shape[:-1] + (5,))
        out[..., 0] = 0.125 + 0.875 * w
        out[..., 1] = 0.3 * w
        out[..., 2] = -0.2 * w
        out[..., 3] = 0.1 * w
        out[..., 4] = 0.1 + 0.9 * w
        return out
    return fn


@pytest.mark.parametrize("dim,p,nx", [(2, 1, [10, 9]), (2, 2, [8, 7]), (2, 3, [8, 8]), (2, 4, [5, 6]), (3, 2, [5, 4, 4]), (3, 3, [4, 4, 3]),
                                      (3, 4, [3, 3, 3])])
def test_rhs_with_active_subcell_fv_blend(dim, p, nx):
    """alpha > 0 in the elements the jump crosses: the troubled-cell correction path (volume scaled by 1 - alpha, interior
    subcell-interface fluxes) against the oracle's subcell_finite_volume_flux restatement."""
    o, tab, h = setup(dim, p, nx, [0.0] * dim, [1.0] * dim, 1.4)
    u = o.project(blast(dim))
    a_ref = o.alpha(u)
    assert (a_ref > 0).sum() >= 2 and (a_ref == 0).sum() >= 1
    want, _ = o.rhs(u)
    l2g = tab["local_to_global"]
    got, alpha = emu.stage(dim, p, u[l2g], tab["face_neighbor"], h, 1.4, mode=1, want_alpha=True)
    assert np.allclose(alpha, a_ref[l2g], rtol=1e-9, atol=1e-12)
    check(got, want[l2g], u[l2g], o, h)


def test_two_species_fields_sources_and_ssprk2_step():
    gamma = 5.0 / 3.0
    dim, p, nx = 2, 3, [6, 5]
    o, tab, h = setup(dim, p, nx, [0.0, 0.0], [1.0, 1.0], gamma, n_species=2, fields=True)
    src = dict(epsilon0=1.3, chi=0.7, charge_over_mass=[0.04, -1.0])
    o.set_sources(True, **src)
    u = o.project(cases.sine_wave(vel=(0.5, 0.3, 0.1), wave=(1, 1, 0)), species=0)
    u = o.project(cases.smooth_blob_3d(0.15), species=1, u=u)
    rng = np.random.default_rng(12345)
    u[:, 10:18, :] = 0.1 * rng.standard_normal(u[:, 10:18, :].shape)
    want, _ = o.rhs(u)
    l2g = tab["local_to_global"]
    ul = u[l2g].copy()
    got = emu.stage(dim, p, ul, tab["face_neighbor"], h, gamma, mode=1, nsp=2, sources=src)
    check(got[:, :10], want[l2g][:, :10], ul[:, :10], o, h)
    assert np.allclose(got[:, 10:], want[l2g][:, 10:], rtol=1e-13, atol=1e-15)
    # one SSPRK2 step (rk.h:97-106) = two launches: f1 = u + dt L(u); u = u/2 + (f1 + dt L(f1))/2, speed of the result fused
    dt = 0.3 * o.recommend_dt(u)
    f1 = emu.stage(dim, p, ul, tab["face_neighbor"], h, gamma, mode=0, dt=dt, a=1.0, beta=0.0, nsp=2, sources=src)
    new, vmax = emu.stage(dim, p, f1, tab["face_neighbor"], h, gamma, mode=0, dt=dt, a=0.5, beta=0.5, dst=ul.copy(), nsp=2, sources=src,
                          want_vmax=True)
    ref = u.copy()
    o.ssprk2_step(ref, dt, 0.0)
    assert (cases.rel_l2_per_component(new, ref[l2g]) < 1e-13).all()
    assert abs(vmax - o.max_transport_speed(ref)) <= 1e-13 * vmax


def test_3d_step_and_low_storage_mode():
    gamma = 1.4
    dim, p, nx = 3, 3, [3, 4, 2]
    o, tab, h = setup(dim, p, nx, [0.0] * 3, [1.0] * 3, gamma)
    u = o.project(cases.smooth_blob_3d(0.1))
    l2g = tab["local_to_global"]
    ul = u[l2g].copy()
    dt = 0.4 * o.recommend_dt(u)
    f1 = emu.stage(dim, p, ul, tab["face_neighbor"], h, gamma, mode=0, dt=dt)
    new, vmax = emu.stage(dim, p, f1, tab["face_neighbor"], h, gamma, mode=0, dt=dt, a=0.5, beta=0.5, dst=ul.copy(), want_vmax=True)
    ref = u.copy()
    o.ssprk2_step(ref, dt, 0.0)
    assert (cases.rel_l2_per_component(new, ref[l2g]) < 1e-13).all()
    assert abs(vmax - o.max_transport_speed(ref)) <= 1e-13 * vmax
    # low-storage stage (mode 2): dst = s + a k, dst2 = s + beta k with k = L(r_in), s = sol_in
    k, _ = o.rhs(u)
    s = 0.5 * ul + 0.01
    d2 = np.zeros_like(ul)
    d1 = emu.stage(dim, p, ul, tab["face_neighbor"], h, gamma, mode=2, a=0.3, beta=0.7, sol_in=s, dst2=d2)
    assert np.allclose(d1, s + 0.3 * k[l2g], rtol=1e-13, atol=1e-13)
    assert np.allclose(d2, s + 0.7 * k[l2g], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize("dim,p,nx", [(2, 3, [6, 8]), (3, 3, [4, 4, 6]), (3, 2, [3, 3, 8])])
def test_sharded_with_ghost_traces_and_split_launches(dim, p, nx):
    """Two ranks: every rank's launch reads the other rank's face traces from the ghost buffer (filled here the way the NCCL
    exchange fills it) and is split into the interface-element launch and the interior launch, as run_stage does."""
    gamma = 1.4
    left, right = [0.0, -5.0, -5.0][:dim], [10.0, 5.0, 5.0][:dim]
    Np, NF = p + 1, (p + 1) ** (dim - 1)
    o = Oracle(dim, p, nx, left, right, gamma=gamma, threads=4)
    h = [(r - l) / n for l, r, n in zip(left, right, nx)]
    u = o.project(cases.isentropic_vortex(gamma))
    want, _ = o.rhs(u)
    for rank in range(2):
        tab = box_tables(dim, nx, [1] * dim, rank=rank, n_ranks=2, group=emu.patch_elems(dim, Np))
        l2g = tab["local_to_global"]
        ghost = np.zeros((tab["n_ghost"], 5, NF))
        for s in range(tab["n_ghost"]):
            ge, side = int(tab["ghost_global_elem"][s]), int(tab["ghost_side"][s])
            nodes = [face_node_to_node(dim, Np, side // 2, side % 2, t) for t in range(NF)]
            ghost[s] = u[ge][:, nodes]
        ul = u[l2g].copy()
        ni = tab["n_interface"]
        assert 0 < ni < len(l2g)
        dst = np.full_like(ul, np.nan)
        emu.stage(dim, p, ul, tab["face_neighbor"], h, gamma, mode=1, dst=dst, ghost=ghost, elem_range=(0, ni))
        emu.stage(dim, p, ul, tab["face_neighbor"], h, gamma, mode=1, dst=dst, ghost=ghost, elem_range=(ni, len(l2g)))
        check(dst, want[l2g], ul, o, h)


# ---- the field system fused into the kernel (perfectly hyperbolic Maxwell fluxes; definition pinned in test_maxwell_cpu.py) ----
MX = dict(light_speed=2.5, chi=0.8, gamma=1.2)
SRC = dict(epsilon0=1.3, chi=0.8, charge_over_mass=[0.04, -1.0])


def _two_fluid(o):
    xyz = o.node_coords()
    x = [xyz[..., d] for d in range(o.dim)] + [0.0, 0.0]
    s = np.sin(2 * np.pi * x[0]) * np.cos(2 * np.pi * x[1]) + 0.3 * np.sin(2 * np.pi * (x[2] + x[0]))
    u = np.zeros(o.shape)
    for sp, (rho0, vel) in enumerate([(25.0, (0.05, -0.02, 0.01)), (1.0, (-0.2, 0.1, 0.05))]):
        prim = np.zeros(xyz.shape[:-1] + (5,))
        prim[..., 0] = rho0 * (1 + 0.1 * s)
        for d in range(3):
            prim[..., 1 + d] = vel[d] * (1 + 0.2 * s)
        prim[..., 4] = 1 + 0.05 * s
        cases.to_state(prim, o.gamma, nc=o.nc, species=sp, u=u)
    for c, amp in enumerate([0.3, -0.2, 0.15, 0.1, -0.05, 0.2, 0.02, -0.03]):
        u[:, 10 + c, :] = 0.1 * amp * (1 + 0.5 * np.cos(2 * np.pi * (x[0] + 0.3 * c)) * np.cos(2 * np.pi * x[1] * (1 + c % 2)))
    return u


@pytest.mark.parametrize("dim,p,nx", [(2, 3, [6, 5]), (2, 2, [5, 4]), (3, 3, [4, 3, 2]), (3, 2, [3, 3, 3]), (2, 4, [3, 3])])
def test_fused_field_system_rhs_step_and_transport_speed(dim, p, nx):
    gamma = 5.0 / 3.0
    o, tab, h = setup(dim, p, nx, [0.0] * dim, [1.0] * dim, gamma, n_species=2, fields=True)
    o.set_sources(True, **SRC)
    o.set_maxwell(True, **MX)
    u = _two_fluid(o)
    l2g = tab["local_to_global"]
    ul = u[l2g].copy()
    want, _ = o.rhs(u)
    got = emu.stage(dim, p, ul, tab["face_neighbor"], h, gamma, mode=1, nsp=2, sources=SRC, maxwell=MX)
    D = oracle.diff_matrix(p + 1)
    scale = cases.summand_scale(ul, gamma, dim, h, D) + cases.field_summand_scale(ul, 2, dim, h, D, **MX)
    err, bound = cases.rhs_error_and_bound(got, want[l2g], scale)
    assert (err <= bound).all(), (err, bound)
    assert (cases.rel_l2_per_component(got, want[l2g])[10:] <= 1e-12).all()
    dt = 0.5 * o.recommend_dt(u)
    f1 = emu.stage(dim, p, ul, tab["face_neighbor"], h, gamma, mode=0, dt=dt, nsp=2, sources=SRC, maxwell=MX)
    new, vmax = emu.stage(dim, p, f1, tab["face_neighbor"], h, gamma, mode=0, dt=dt, a=0.5, beta=0.5, dst=ul.copy(), nsp=2, sources=SRC,
                          maxwell=MX, want_vmax=True)
    ref = u.copy()
    o.ssprk2_step(ref, dt, 0.0)
    assert (cases.rel_l2_per_component(new, ref[l2g]) < 1e-13).all()
    assert abs(vmax - o.max_transport_speed(ref)) <= 1e-13 * vmax


def test_fused_field_system_sharded():
    """the ghost traces carry all nc components once the field system is evolved"""
    dim, p, nx, gamma = 3, 3, [4, 4, 6], 5.0 / 3.0
    Np, NF = p + 1, (p + 1) ** (dim - 1)
    o = Oracle(dim, p, nx, [0.0] * 3, [1.0] * 3, gamma=gamma, n_species=2, fields_enabled=True, threads=4)
    o.set_sources(True, **SRC)
    o.set_maxwell(True, **MX)
    h = [1.0 / n for n in nx]
    u = _two_fluid(o)
    want, _ = o.rhs(u)
    for rank in range(2):
        tab = box_tables(dim, nx, [1] * dim, rank=rank, n_ranks=2, group=emu.patch_elems(dim, Np))
        l2g = tab["local_to_global"]
        ghost = np.zeros((tab["n_ghost"], 18, NF))
        for s in range(tab["n_ghost"]):
            ge, side = int(tab["ghost_global_elem"][s]), int(tab["ghost_side"][s])
            ghost[s] = u[ge][:, [face_node_to_node(dim, Np, side // 2, side % 2, t) for t in range(NF)]]
        ul = u[l2g].copy()
        got = emu.stage(dim, p, ul, tab["face_neighbor"], h, gamma, mode=1, nsp=2, ghost=ghost, sources=SRC, maxwell=MX)
        D = oracle.diff_matrix(Np)
        scale = cases.summand_scale(ul, gamma, dim, h, D) + cases.field_summand_scale(ul, 2, dim, h, D, **MX)
        err, bound = cases.rhs_error_and_bound(got, want[l2g], scale)
        assert (err <= bound).all(), (rank, err, bound)


@pytest.mark.parametrize("dim,p,nx", [(3, 3, [4, 3, 2]), (3, 2, [3, 3, 3]), (3, 4, [2, 2, 3]), (2, 3, [6, 5]), (2, 1, [7, 6]), (3, 1, [4, 3, 3]),
                                      (3, 5, [2, 2, 2])])
def test_stand_alone_field_kernel_rhs_step_and_transport_speed(dim, p, nx):
    """The 3-D field system's kernel (dgsem_maxwell_thread.cuh) on the host: field components of the RHS, a full SSPRK2 step
    next to the pencil kernel that skips the fields, and the field system's share of the transport speed."""
    gamma, nsp, nf = 5.0 / 3.0, 2, 10
    o, tab, h = setup(dim, p, nx, [0.0] * dim, [1.0] * dim, gamma, n_species=nsp, fields=True)
    o.set_sources(True, **SRC)
    o.set_maxwell(True, **MX)
    u = _two_fluid(o)
    l2g = tab["local_to_global"]
    ul = u[l2g].copy()
    want, _ = o.rhs(u)
    kw = dict(nsp=nsp, sources=SRC)
    got = emu.stage(dim, p, ul, tab["face_neighbor"], h, gamma, mode=1, maxwell=MX, field_kernel=True, **kw)
    assert (cases.rel_l2_per_component(got, want[l2g])[nf:] <= 1e-12).all()
    assert np.abs(got[:, :nf, :]).max() == 0.0                      # the fluid components are not this kernel's business
    if not 3 <= p + 1 <= 5:
        return                                                      # (the pencil kernel serves Np = 3..5: RHS check only)
    # SSPRK2 with both kernels, stage by stage: fluids by the pencil kernel (which then leaves the fields alone), fields by the
    # field kernel
    dt = 0.5 * o.recommend_dt(u)

    def stage(src_vec, dst, a, beta):
        dst, v_fluid = emu.stage(dim, p, src_vec, tab["face_neighbor"], h, gamma, mode=0, dt=dt, a=a, beta=beta, dst=dst, want_vmax=True,
                                 fields_skip=True, **kw)
        dst, v_field = emu.stage(dim, p, src_vec, tab["face_neighbor"], h, gamma, mode=0, dt=dt, a=a, beta=beta, dst=dst, want_vmax=True,
                                 maxwell=MX, field_kernel=True, **kw)
        return dst, max(v_fluid, v_field)
    f1, _ = stage(ul, np.zeros_like(ul), 1.0, 0.0)
    new, vmax = stage(f1, ul.copy(), 0.5, 0.5)
    ref = u.copy()
    o.ssprk2_step(ref, dt, 0.0)
    assert (cases.rel_l2_per_component(new, ref[l2g]) < 1e-13).all()
    assert abs(vmax - o.max_transport_speed(ref)) <= 1e-13 * vmax


def test_stand_alone_field_kernel_sharded():
    """the field kernel reads the other rank's field traces from the ghost buffer (all nc components per face node)"""
    dim, p, nx, gamma = 3, 3, [4, 4, 6], 5.0 / 3.0
    Np, NF = p + 1, (p + 1) ** (dim - 1)
    o = Oracle(dim, p, nx, [0.0] * 3, [1.0] * 3, gamma=gamma, n_species=2, fields_enabled=True, threads=4)
    o.set_sources(True, **SRC)
    o.set_maxwell(True, **MX)
    h = [1.0 / n for n in nx]
    u = _two_fluid(o)
    want, _ = o.rhs(u)
    for rank in range(2):
        tab = box_tables(dim, nx, [1] * dim, rank=rank, n_ranks=2, group=emu.patch_elems(dim, Np))
        l2g = tab["local_to_global"]
        ghost = np.zeros((tab["n_ghost"], 18, NF))
        for s in range(tab["n_ghost"]):
            ge, side = int(tab["ghost_global_elem"][s]), int(tab["ghost_side"][s])
            ghost[s] = u[ge][:, [face_node_to_node(dim, Np, side // 2, side % 2, t) for t in range(NF)]]
        ul = u[l2g].copy()
        # split launches like the product's: interface elements, then the interior ones
        n_if = int(tab["n_interface"]) if "n_interface" in tab else 0
        got = np.zeros_like(ul)
        for rng in ([(0, n_if), (n_if, ul.shape[0])] if 0 < n_if < ul.shape[0] else [(0, ul.shape[0])]):
            emu.stage(dim, p, ul, tab["face_neighbor"], h, gamma, mode=1, nsp=2, ghost=ghost, sources=SRC, maxwell=MX, field_kernel=True,
                      dst=got, elem_range=rng)
        assert (cases.rel_l2_per_component(got, want[l2g])[10:] <= 1e-12).all(), rank


def _pslot(n):
    # dgsem_pencil_stage.cuh::pslot for Np = 4
    return n ^ ((n >> 3) & 1) ^ (((n >> 4) & 1) * 6)


def _pencil_node(dim, d, pe, m, np_=4):
    # dgsem_pencil_stage.cuh::pencil_node
    if dim == 2:
        return pe * np_ + m if d == 0 else pe + np_ * m
    if d == 0:
        return pe * np_ + m
    if d == 1:
        return (pe % np_) + np_ * m + np_ * np_ * (pe // np_)
    return pe + np_ * np_ * m


@pytest.mark.parametrize("dim", [2, 3])
def test_swizzled_planes_are_bijective_and_bank_conflict_free(dim):
    """The Np = 4 layout claim of DESIGN.md section 3: a 16-byte access of a quarter-warp (8 consecutive pencil owners, the
    unit the shared-memory pipe serves per wavefront) touches 8 distinct 16-byte bank groups, whatever direction the
    threads own and whichever node m of their pencil they address; and the swizzle is a permutation of the patch's nodes."""
    np_, nn, npen = 4, 4 ** dim, 4 ** (dim - 1)
    elems = emu.patch_elems(dim, np_)
    nodes = elems * nn
    assert sorted(_pslot(n) for n in range(nodes)) == list(range(nodes))
    for d in range(dim):
        for m in range(np_):
            for q0 in range(0, elems * npen, 8):
                groups = set()
                for tid in range(q0, q0 + 8):
                    le, pe = divmod(tid, npen)
                    groups.add(_pslot(le * nn + _pencil_node(dim, d, pe, m)) % 8)
                assert len(groups) == 8, (dim, d, m, q0)
