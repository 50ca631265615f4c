"""Small run of every kernel added for general geometry, low-storage RK and the streamed host step, meant to be executed
under compute-sanitizer (memcheck / racecheck) on the GPU box:
    compute-sanitizer --tool racecheck python scripts/sanitize_general.py"""
import os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mesh_cases as mc
from warpii_b200 import BC_INFLOW, BC_OUTFLOW, BC_WALL
from warpii_b200.capi import MeshSolver, mapped_metrics

G = 1.4
for dim, p, nx, periodic in [(2, 3, [5, 3], False), (3, 2, [3, 2, 3], True), (2, 4, [3, 3], True), (3, 3, [2, 3, 2], False), (1, 2, [9], False)]:
    left, right = [0.0] * dim, [1.0, 1.2, 0.9][:dim]
    per = [int(periodic)] * dim
    bc = None if periodic else np.array([[[BC_OUTFLOW, BC_WALL, BC_INFLOW][f % 3] for f in range(2 * dim)]])
    mesh, xyz = mc.mapped_box(dim, p, nx, left, right, per, mc.wavy(left, right, 0.03))
    geo = mapped_metrics(dim, p, xyz, mesh["face_neighbor"], None, mesh["bf_elem"], mesh["bf_side"])
    g = MeshSolver(dim, p, mesh, geo, n_boundaries=0 if periodic else 2 * dim, bc_kinds=bc, gamma=G, n_vectors=4)
    if not periodic:
        for f in range(2 * dim):
            g.set_inflow(0, f, mc.to_conserved(np.array([1.05, 0.3, 0.0, 0.0, 1.0]), G))
    prim = mc.periodic_state(G, left, right, dim)(mc.box_node_coords(dim, p, nx, left, right))
    mc.add_kinks(prim)
    u = mc.state_from(prim, G)
    g.upload(0, u)
    g.rhs(1, 0)
    dt = g.recommend_dt(0)
    g.ssprk2_step(dt, 0.0)
    g.advance_to(0.0, 1e9, max_steps=3)
    g.lsrk_stage(2, 3, 0, 0, 0.1 * dt, 0.2 * dt)
    g.lsrk_stage(2, 1, 2, 3, 0.1 * dt, 0.0)
    g.global_integral(0)
    g.shock_indicator(0)
    host = g.download(0)
    g.host_step(host, host, dt, 0.0, n_slabs=3)
    assert np.isfinite(host).all()
    g.close()
    print("ok", dim, p, nx, flush=True)
verts, cells = mc.hexagon_blocks(2)
mesh, xyz = mc.quad_mesh(verts, cells, 3, mc.hexagon_boundary_id, warp=mc.swirl_warp(0.03))
geo = mapped_metrics(2, 3, xyz, mesh["face_neighbor"], mesh["neighbor_face"], mesh["bf_elem"], mesh["bf_side"])
g = MeshSolver(2, 3, mesh, geo, n_boundaries=3, bc_kinds=np.array([[BC_WALL, BC_INFLOW, BC_OUTFLOW]]), gamma=G)
g.set_inflow(0, 1, mc.to_conserved(np.array([1.1, 0.3, 0.2, 0.0, 1.0]), G))
prim = mc.smooth_state(G, 2)(xyz)
mc.add_kinks(prim)
g.upload(0, mc.state_from(prim, G))
g.advance_to(0.0, 1e9, max_steps=3)
assert np.isfinite(g.download(0)).all()
g.close()
print("ok hexagon", flush=True)
