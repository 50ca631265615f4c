"""Worker of tests/test_gpu_multi.py: one rank per GPU (torch.distributed.run), element-sharded run against the
global CPU oracle.  Exits non-zero on any mismatch."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dgsem_cases as cases  # noqa: E402
from oracle import Oracle  # noqa: E402
from warpii_b200 import BoxSolver, nccl_unique_id  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def fresh_id():
        # one NCCL unique id per communicator: created on rank 0 through the ABI, broadcast by the caller
        t = torch.tensor(list(nccl_unique_id()) if rank == 0 else [0] * 128, dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())

    for (dim, p, nx, left, right, ic, gamma) in [
        (2, 3, [8, 12], [0.0, -5.0], [10.0, 5.0], cases.isentropic_vortex(1.4), 1.4),
        (3, 2, [4, 4, 6], [0.0, 0.0, 0.0], [1.0, 1.0, 1.0], cases.smooth_blob_3d(0.1), 5.0 / 3.0),
        (1, 3, [16], [0.0], [1.0], cases.sine_wave(), 5.0 / 3.0),
    ]:
        o = Oracle(dim, p, nx, left, right, gamma=gamma, threads=4)
        u = o.project(ic)
        g = BoxSolver(dim, p, nx, left, right, gamma=gamma, rank=rank, n_ranks=world, device=local)
        g.attach_comm(fresh_id())
        assert g.n_elems < o.n_elems
        # one RHS with halo exchange
        g.upload_global(0, u)
        g.rhs(1, 0)
        got = g.download(1)
        want, _ = o.rhs(u)
        err = cases.rel_l2_per_component(got, want[g.l2g])
        assert (err <= 1e-12).all(), (rank, dim, err)
        # dt: max over ranks (NCCL all-reduce) equals the global oracle value
        dt = g.recommend_dt(0)
        assert abs(dt - o.recommend_dt(u)) <= 1e-13 * dt, (dt, o.recommend_dt(u))
        # global integrals are summed over ranks
        assert np.allclose(g.global_integral(0), o.global_integral(u), rtol=1e-13, atol=1e-13)
        # 20 steps: every rank takes the same dt sequence; conservation and parity
        ic_int = o.global_integral(u)
        t, steps = g.advance_to(0.0, 1e9, max_steps=20)
        o.solve(u, t, max_steps=20)
        got = g.download(0)
        err = cases.rel_l2_per_component(got, u[g.l2g])
        assert (err <= 1e-11).all(), (rank, dim, err)
        now = g.global_integral(0)
        assert np.allclose(now, ic_int, rtol=1e-12, atol=1e-12)
        # the sharded result is bit-identical to the single-GPU result of the same library
        if rank == 0:
            g1 = BoxSolver(dim, p, nx, left, right, gamma=gamma, device=local)
            g1.upload_global(0, o.project(ic))
            g1.advance_to(0.0, 1e9, max_steps=20)
            single = g1.download_global(0)
            assert np.array_equal(single[g.l2g], got), "sharded and single-GPU runs differ"
            g1.close()
        g.close()
        dist.barrier()
    # ---- streamed host step on a sharded context (interface elements + exchanges on the communication stream, interior
    # slabs on the main stream) == upload + SSPRK2 step + download, bit for bit ---------------------------------------------------
    for (dim, p, nx, slabs) in [(2, 3, [16, 24], 4), (3, 3, [6, 6, 12], 0), (3, 2, [4, 4, 8], 3)]:
        left, right = [0.0, -5.0, -5.0][:dim], [10.0, 5.0, 5.0][:dim]
        o = Oracle(dim, p, nx, left, right, gamma=1.4, threads=4)
        u = o.project(cases.isentropic_vortex(1.4))
        g = BoxSolver(dim, p, nx, left, right, gamma=1.4, rank=rank, n_ranks=world, device=local)
        g.attach_comm(fresh_id())
        ul = np.ascontiguousarray(u[g.l2g])
        g.upload(0, ul)
        dt = g.recommend_dt(0)
        want, dts = ul.copy(), []
        for k in range(3):
            g.upload(0, want)
            g.ssprk2_step(dt if k == 0 else dts[-1], 0.0)
            want = g.download(0)
            dts.append(g.recommend_dt(0))
        host = torch.empty(ul.size, dtype=torch.float64).pin_memory().numpy().reshape(ul.shape)
        host[...] = ul
        got, d = [], dt
        for k in range(3):
            d = g.host_step(host, host, d, 0.0, n_slabs=slabs)
            got.append(d)
        assert np.array_equal(host, want), (rank, dim, "streamed sharded host step differs from the plain sequence")
        assert got == dts, (got, dts)
        g.close()
        dist.barrier()
    # ---- two fluids + the field system (PHM Maxwell fluxes + sources): the halo carries the field traces too ----------------
    from test_gpu_maxwell import MX, SRC, two_fluid_state
    for (dim, p, nx) in [(2, 3, [6, 8]), (3, 3, [4, 4, 6]), (1, 2, [12])]:
        left, right = [0.0] * dim, [1.0] * dim
        kw = dict(gamma=5.0 / 3.0, n_species=2, fields_enabled=True)
        o = Oracle(dim, p, nx, left, right, threads=4, **kw)
        o.set_sources(True, SRC["epsilon0"], SRC["chi"], SRC["charge_over_mass"])
        o.set_maxwell(True, **MX)
        u = two_fluid_state(o)
        g = BoxSolver(dim, p, nx, left, right, rank=rank, n_ranks=world, device=local, **kw)
        g.set_sources(True, SRC["epsilon0"], SRC["chi"], SRC["charge_over_mass"])
        g.set_maxwell(True, **MX)
        g.attach_comm(fresh_id())
        g.upload_global(0, u)
        g.rhs(1, 0)
        want, _ = o.rhs(u)
        err = cases.rel_l2_per_component(g.download(1), want[g.l2g])
        assert (err <= 1e-12).all(), (rank, dim, err)
        dt = g.recommend_dt(0)
        assert abs(dt - o.recommend_dt(u)) <= 1e-13 * dt
        t, steps = g.advance_to(0.0, 1e9, max_steps=20)
        o.solve(u, t, max_steps=20)
        got = g.download(0)
        assert (cases.rel_l2_per_component(got, u[g.l2g]) <= 1e-11).all()
        if rank == 0:
            g1 = BoxSolver(dim, p, nx, left, right, device=local, **kw)
            g1.set_sources(True, SRC["epsilon0"], SRC["chi"], SRC["charge_over_mass"])
            g1.set_maxwell(True, **MX)
            g1.upload_global(0, two_fluid_state(o))
            g1.advance_to(0.0, 1e9, max_steps=20)
            assert np.array_equal(g1.download_global(0)[g.l2g], got), "sharded and single-GPU runs differ (field system)"
            g1.close()
        g.close()
        dist.barrier()
    # ---- curved elements (general-geometry kernels), slab-sharded: halo traces, ghost-face normals from the own element ----
    import mesh_cases as mc
    from oracle import GeneralOracle
    from warpii_b200.capi import mapped_metrics
    for (dim, p, nx) in [(2, 3, [8, 12]), (3, 2, [4, 4, 6])]:
        left, right = [0.0] * dim, [1.0, 1.2, 0.9][:dim]
        warp = mc.wavy(left, right, 0.04)
        mesh, xyz = mc.mapped_box(dim, p, nx, left, right, [1] * dim, warp)
        o = GeneralOracle(dim, p, mesh, mapped_metrics(dim, p, xyz, mesh["face_neighbor"]), gamma=1.4, threads=4)
        g = BoxSolver.mapped(dim, p, nx, left, right, lambda x: warp(x[None, :])[0], gamma=1.4, rank=rank, n_ranks=world, device=local)
        g.attach_comm(fresh_id())
        assert g.n_elems < o.n_elems
        assert np.abs(g.node_coords() - xyz[g.l2g]).max() <= 1e-14
        prim = mc.periodic_state(1.4, left, right, dim)(mc.box_node_coords(dim, p, nx, left, right))
        mc.add_kinks(prim)
        u = mc.state_from(prim, 1.4)
        g.upload_global(0, u)
        g.rhs(1, 0)
        want, _ = o.rhs(u)
        err = cases.rel_l2_per_component(g.download(1), want[g.l2g])
        assert (err <= 1e-12).all(), (rank, dim, err)
        dt = g.recommend_dt(0)
        assert abs(dt - o.recommend_dt(u)) <= 1e-13 * dt
        assert np.allclose(g.global_integral(0), o.global_integral(u), rtol=1e-13, atol=1e-13)
        ic_int = o.global_integral(u)
        t, steps = g.advance_to(0.0, 1e9, max_steps=20)
        o.solve(u, t, max_steps=20)
        err = cases.rel_l2_per_component(g.download(0), u[g.l2g])
        assert (err <= 1e-10).all(), (rank, dim, err)
        now = g.global_integral(0)
        assert abs(now[0] - ic_int[0]) <= 1e-12 * abs(ic_int[0]) and abs(now[4] - ic_int[4]) <= 1e-12 * abs(ic_int[4])
        if dim == 2:
            assert np.allclose(now, ic_int, rtol=1e-12, atol=1e-12)
        g.close()
        dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank}: multi-GPU parity ok")


if __name__ == "__main__":
    main()
