#!/usr/bin/env python
"""Plain per-component relative L2 error of one RHS evaluation, CUDA path (through the C ABI) against the CPU oracle,
on the parity-test meshes AND at BASELINE.json's sizes.  Written to gpurun_out/parity_<round>.json (copied to
profiles/ once looked at).  usage (GPU box): python tests/tools/parity_table.py [round] [--quick]

"plain" = ||got - want||_2 / ||want||_2 per component, no guard; a component is only excused when the oracle's own
RHS norm is below 1e-9 x the magnitude of the terms that are differenced to form it (pure cancellation noise in
BOTH codes); those entries are listed under "noise_components"."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dgsem_cases as cases  # noqa: E402
import oracle  # noqa: E402
from oracle import Oracle  # noqa: E402
from warpii_b200 import BoxSolver  # noqa: E402

sys.path.insert(0, ROOT)
import bench  # noqa: E402  (workload table + initial conditions)


def one_case(name, dim, p, nx, left, right, gamma, make_u, n_species=1, fields=False, periodic=None, bc=None, sources=None,
             threads=None, maxwell=None):
    threads = threads or (os.cpu_count() or 1)
    kw = dict(gamma=gamma, n_species=n_species, fields_enabled=fields)
    if periodic is not None:
        kw.update(periodic=periodic, bc_kinds=bc)
    o = Oracle(dim, p, nx, left, right, threads=threads, **kw)
    gkw = dict(gamma=gamma, n_species=n_species, fields_enabled=fields)
    if periodic is not None:
        gkw.update(periodic=periodic, bc_kinds=bc, n_boundaries=2 * dim if not all(periodic) else None)
    g = BoxSolver(dim, p, nx, left, right, **gkw)
    if sources:
        o.set_sources(True, **sources)
        g.set_sources(True, **sources)
    if maxwell:
        o.set_maxwell(True, **maxwell)
        g.set_maxwell(True, **maxwell)
    u = make_u(o)
    t0 = time.perf_counter()
    want, _ = o.rhs(u)
    t_o = time.perf_counter() - t0
    # CPU-vs-CPU rounding floor: the same source compiled with multiply-add contraction (logarithm and pressure pinned)
    of = Oracle(dim, p, nx, left, right, threads=threads, variant="fma", **kw)
    if sources:
        of.set_sources(True, **sources)
    if maxwell:
        of.set_maxwell(True, **maxwell)
    want_f, _ = of.rhs(u)
    del of
    g.upload_global(0, u)
    g.rhs(1, 0)
    got = g.download_global(1)
    g.close()
    h = [(r - l) / n for l, r, n in zip(left, right, nx)]
    scale = cases.summand_scale(u, gamma, dim, h, oracle.diff_matrix(p + 1))
    if maxwell:
        scale = scale + cases.field_summand_scale(u, n_species, dim, h, oracle.diff_matrix(p + 1), **maxwell)
    nfl = 5 * n_species + (8 if maxwell else 0)
    plain, noise, floor, ulps = [], [], [], []
    for c in range(want.shape[1]):
        den = np.linalg.norm(want[:, c, :])
        num = np.linalg.norm(got[:, c, :] - want[:, c, :])
        numf = np.linalg.norm(want_f[:, c, :] - want[:, c, :])
        ulps.append(float(num / (2.0 ** -52 * scale[c])) if scale[c] > 0 else None)
        if c < nfl and den < 1e-9 * scale[c]:
            noise.append(c)
            plain.append(None)
            floor.append(None)
        else:
            plain.append(float(num / den) if den > 0 else float(num))
            floor.append(float(numf / den) if den > 0 else float(numf))
    worst = max([v for v in plain if v is not None] or [0.0])
    err, bound = cases.rhs_error_and_bound(got, want, scale)
    rec = {"case": name, "dim": dim, "p": p, "nx": list(nx), "n_dofs": int(u.size), "rel_l2_per_component": plain,
           "cpu_vs_cpu_floor_per_component": floor, "error_in_ulps_of_differenced_terms": ulps,
           "passes_test_criterion": bool((err <= bound).all()), "kappa": [float(s / w) if w > 0 else None for s, w in
                                                                       zip(scale, [np.linalg.norm(want[:, c, :]) for c in range(want.shape[1])])],
           "noise_components": noise, "worst": worst, "oracle_seconds": round(t_o, 2),
           "want_norm": [float(np.linalg.norm(want[:, c, :])) for c in range(want.shape[1])],
           "summand_scale": [float(s) for s in scale]}
    print(json.dumps({k: rec[k] for k in ("case", "n_dofs", "worst", "passes_test_criterion", "noise_components", "oracle_seconds")}), flush=True)
    return rec


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "r02"
    quick = "--quick" in sys.argv
    out = []
    ic = lambda fn: (lambda o: o.project(fn))
    small = [
        ("1d_p2_sine", 1, 2, [20], [0.0], [1.0], 5 / 3, ic(cases.sine_wave())),
        ("2d_p3_vortex_16", 2, 3, [16, 16], [0.0, -5.0], [10.0, 5.0], 1.4, ic(cases.isentropic_vortex())),
        ("2d_p5_blob", 2, 5, [4, 6], [0.0, 0.0], [1.0, 1.0], 5 / 3, ic(cases.smooth_blob_3d())),
        ("3d_p3_blob", 3, 3, [5, 4, 6], [0.0] * 3, [1.0] * 3, 5 / 3, ic(cases.smooth_blob_3d())),
        ("3d_p4_vortex_4", 3, 4, [4, 4, 4], [0.0, -5.0, -5.0], [10.0, 5.0, 5.0], 1.4, ic(cases.isentropic_vortex())),
        ("3d_p2_blob", 3, 2, [6, 6, 3], [0.0] * 3, [1.0, 2.0, 1.0], 1.4, ic(cases.smooth_blob_3d(0.1))),
        ("2d_p3_vortex_128", 2, 3, [128, 128], [0.0, -5.0], [10.0, 5.0], 1.4, ic(cases.isentropic_vortex())),
        ("2d_p3_vortex_128_quarter", 2, 3, [128, 128], [2.5, -1.25], [5.0, 1.25], 1.4, ic(cases.isentropic_vortex())),
        ("3d_p4_vortex_8_fine", 3, 4, [8, 8, 8], [4.0, -0.5, -0.5], [5.0, 0.5, 0.5], 1.4, ic(cases.isentropic_vortex())),
        ("3d_p3_vortex_16", 3, 3, [16, 16, 16], [0.0, -5.0, -5.0], [10.0, 5.0, 5.0], 1.4, ic(cases.isentropic_vortex())),
    ]
    for c in small:
        out.append(one_case(*c))

    # BASELINE sizes (bench.py's workload table; the same initial conditions the bench times)
    def wl(name, nx=None, left=None, right=None):
        w = dict(bench.WORKLOADS[name])
        if nx is not None:
            w["nx"], w["left"], w["right"] = nx, left, right
        mk = lambda o: bench.build_ic(w, o.node_coords())
        return one_case(name if nx is None else f"{name}_{'x'.join(map(str, nx))}", w["dim"], w["p"], w["nx"], w["left"], w["right"],
                        w["gamma"], mk, n_species=w.get("n_species", 1), fields=w.get("fields", False),
                        periodic=w.get("periodic"), bc=w.get("bc"), sources=w.get("sources"), maxwell=w.get("maxwell"))
    out.append(wl("C2"))
    if not quick:
        out.append(wl("C3"))
        out.append(wl("C3s"))
        # C4 shard: 32^3 elements of the 128^3 mesh (same h, degree 4)
        out.append(wl("C4s", [32, 32, 32], [3.75, -1.25, -1.25], [6.25, 1.25, 1.25]))
        out.append(wl("C5s"))
        out.append(wl("V3D3"))
        out.append(wl("N3D", [48, 48, 48], [0.0, -5.0, -5.0], [15.0, 2.5, 2.5]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    path = os.path.join(ROOT, "gpurun_out", f"parity_{rnd}.json")
    with open(path, "w") as f:
        json.dump({"tolerance": 1e-12, "metric": "plain relative L2 per component, one RHS evaluation, GPU vs oracle; cpu_vs_cpu_floor = the oracle source "
                   "compiled with multiply-add contraction vs the oracle (two legitimate FP64 builds of one source); ulps = GPU error in "
                   "units of 2^-52 x the magnitude of the terms differenced to form the RHS (the test criterion allows 2)",
                   "cases": out, "worst_overall": max(c["worst"] for c in out)}, f, indent=1)
    print("worst overall", max(c["worst"] for c in out), "->", path)


if __name__ == "__main__":
    main()
