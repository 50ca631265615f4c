"""Stage-kernel time of the general-geometry path against the Cartesian path on a workload of the same size:
a doubly periodic [0,10]x[-5,5] box, degree 3, n x n elements (n = 512: BASELINE config 2), isentropic vortex; the general
run maps the box through a smooth periodic perturbation (curved elements, metric terms read from HBM).
Usage: python scripts/general_rate.py [n] [steps]"""
import json, os, sys
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dgsem_cases as cases
import mesh_cases as mc
from warpii_b200 import BoxSolver, box_tables, elems_per_block
from warpii_b200.capi import MeshSolver, mapped_metrics

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dim, p, gamma = 2, 3, 1.4
left, right = [0.0, -5.0], [10.0, 5.0]
ic = cases.isentropic_vortex(gamma)
out = {}

def run(g, label, n_dofs):
    g.advance_to(0.0, 1e30, max_steps=4)
    g.stage_timing(True)
    g.advance_to(0.0, 1e30, max_steps=steps)
    ms, launches = g.stage_timing(False)
    per_stage = ms / launches
    out[label] = {"stage_ms": per_stage, "dof_updates_per_s_stage_only": n_dofs / (per_stage * 1e-3), "launches": launches}
    print(label, out[label], flush=True)

# Cartesian path
g = BoxSolver(dim, p, [n, n], left, right, gamma=gamma)
g.set_state(cases.to_state(ic(g.node_coords()), gamma))
run(g, "cartesian", g.n_dofs)
l2g = g.l2g.copy()
g.close()

# general path on the same element numbering (patches of elems_per_block elements), curved mapping
t = box_tables(dim, [n, n], [1, 1], group=elems_per_block(dim, p))
assert np.array_equal(t["local_to_global"], l2g)
mesh = {"face_neighbor": t["face_neighbor"], "neighbor_face": None, "bf_elem": [], "bf_side": [], "bf_id": []}
ref = mc.ref_nodes(dim, p)
ex, ey = l2g % n, l2g // n
h = [(right[d] - left[d]) / n for d in range(dim)]
box_xyz = np.zeros((n * n, ref.shape[0], dim))
box_xyz[:, :, 0] = left[0] + (ex[:, None] + ref[None, :, 0]) * h[0]
box_xyz[:, :, 1] = left[1] + (ey[:, None] + ref[None, :, 1]) * h[1]
for label, amp in (("general_identity", 0.0), ("general_curved", 0.02)):
    xyz = mc.wavy(left, right, amp)(box_xyz) if amp else box_xyz
    geo = mapped_metrics(dim, p, xyz, mesh["face_neighbor"])
    g = MeshSolver(dim, p, mesh, geo, gamma=gamma)
    g.upload(0, cases.to_state(ic(box_xyz), gamma))
    run(g, label, g.n_dofs)
    g.close()

# ---- 3D: 48^3 degree-3 elements, Cartesian kernel vs general kernel on the wavy box ------------------------------------
n3 = int(sys.argv[3]) if len(sys.argv) > 3 else 48
if n3 > 0:
    dim, left, right = 3, [0.0, -5.0, -5.0], [10.0, 5.0, 5.0]
    g = BoxSolver(dim, p, [n3] * 3, left, right, gamma=gamma)
    g.set_state(cases.to_state(ic(g.node_coords()), gamma))
    run(g, "cartesian_3d", g.n_dofs)
    l2g = g.l2g.copy()
    box_xyz = g.node_coords()
    g.close()
    t = box_tables(dim, [n3] * 3, [1, 1, 1], group=elems_per_block(dim, p))
    assert np.array_equal(t["local_to_global"], l2g)
    mesh = {"face_neighbor": t["face_neighbor"], "neighbor_face": None, "bf_elem": [], "bf_side": [], "bf_id": []}
    geo = mapped_metrics(dim, p, mc.wavy(left, right, 0.01)(box_xyz), mesh["face_neighbor"])
    g = MeshSolver(dim, p, mesh, geo, gamma=gamma)
    del geo
    g.upload(0, cases.to_state(ic(box_xyz), gamma))
    run(g, "general_curved_3d", g.n_dofs)
    g.close()

# ---- low-storage RK: one fused stage (mode 2) on the C2 box -----------------------------------------------------------
dim, left, right = 2, [0.0, -5.0], [10.0, 5.0]
g = BoxSolver(dim, p, [n, n], left, right, gamma=gamma)
g.set_state(cases.to_state(ic(g.node_coords()), gamma))
dt = g.recommend_dt(0)
g.lsrk_stage(2, 1, 0, 0, 0.2 * dt, 0.1 * dt)
g.stage_timing(True)
for k in range(10):
    g.lsrk_stage(2, 3 if k % 2 == 0 else 1, 2, 1 if k % 2 == 0 else 3, 0.02 * dt, 0.01 * dt)
ms, launches = g.stage_timing(False)
out["lsrk_stage_cartesian"] = {"stage_ms": ms / launches, "launches": launches, "bytes_per_dof": 32}
print("lsrk_stage_cartesian", out["lsrk_stage_cartesian"], flush=True)
g.close()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "general_rate.json"), "w"), indent=1)
