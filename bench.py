#!/usr/bin/env python
"""bench.py -- FP64 DoF-updates/s of the ES-DGSEM RHS + SSPRK2 hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm  (libwarpii_b200.so on N GPUs, one rank per GPU)
  python bench.py --impl reference --steps K --warmup W    reference arm: the CPU restatement of the reference
                                                           algorithm (oracle/), all host threads, rank 0 only

One "step" is one SSPRK2 time step of the workload = 2 RHS evaluations (+ the CFL reduction the time loop asks
for every step, fluid_flux_es_dgsem_operator.h:442-448), so DoF-updates per step = 2 * n_dofs.
Default workload: the shape BASELINE.json quotes its target on, "degree-3 3D five-moment RHS" (N3D: 3D, p=3, two species +
8 field components, 64^3 elements per GPU, periodic).  For N>1 the mesh is extended by 64 element layers per GPU (weak
scaling) and sharded in slabs with an NCCL halo exchange per RHS (a 64x64-face 3-D halo, 2.9 MB per direction and RHS).
At N=1 the line also carries `other_workloads`: BASELINE configs 2-5 at their stated shapes (C2, C3, C4s = one GPU's share
of config 4, C5s) measured in the same run.  For N>1 a sharded-vs-oracle RHS check runs outside the timed region and is
reported as `parity_check`.  The timed region is repeated 5 times (K steps each); `ms_per_step` is the median repetition.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import dgsem_cases as cases  # noqa: E402

METRIC = "FP64 DoF-updates/sec (RHS evals x DoFs/s)"
UNIT = "DoF-updates/s"
BYTES_PER_DOF_UPDATE = 20.0   # stage 1 = 16 B, stage 2 = 24 B per DoF, SSPRK2 average 20 B

WORKLOADS = {
    # name: (dim, p, nx per GPU, left, right, gamma, ic, n_species, fields)
    "C2": dict(dim=2, p=3, nx=[512, 512], left=[0.0, -5.0], right=[10.0, 5.0], gamma=1.4, ic="vortex",
               label="C2: 2D Euler periodic isentropic vortex, degree 3, 512x512 elements"),
    "C3p": dict(dim=2, p=3, nx=[1024, 1024], left=[0.0, -5.0], right=[10.0, 5.0], gamma=1.4, ic="vortex",
                label="2D Euler periodic vortex, degree 3, 1024x1024 elements (C3 size, doubly periodic)"),
    "V3D3": dict(dim=3, p=3, nx=[64, 64, 64], left=[0.0, -5.0, -5.0], right=[10.0, 5.0, 5.0], gamma=1.4, ic="vortex",
                 label="3D Euler periodic vortex, degree 3, 64^3 elements"),
    "C3": dict(dim=2, p=3, nx=[1024, 1024], left=[-0.5, 0.0], right=[0.5, 2 * np.pi / (1.2 * np.pi)], gamma=5.0 / 3.0, ic="kh",
               periodic=[0, 1], bc=[[0, 0, 0, 0]],
               label="C3: 2D Euler Kelvin-Helmholtz, walls in x, periodic in y, degree 3, 1024x1024 elements"),
    "C3s": dict(dim=2, p=3, nx=[1024, 1024], left=[-0.5, 0.0], right=[0.5, 2 * np.pi / (1.2 * np.pi)], gamma=5.0 / 3.0, ic="kh_steep",
                periodic=[0, 1], bc=[[0, 0, 0, 0]],
                label="C3 with an under-resolved interface (tanh width = one element): subcell-FV blend active, walls in x, "
                      "degree 3, 1024x1024 elements"),
    "C5s": dict(dim=2, p=3, nx=[1024, 512], left=[0.0, -5.0], right=[20.0, 5.0], gamma=5.0 / 3.0, ic="two_fluid", n_species=2,
                fields=True, sources=dict(epsilon0=1.0, chi=1.0, charge_over_mass=[1.0 / 25.0, -1.0]),
                maxwell=dict(light_speed=10.0, chi=1.0, gamma=1.0),
                label="C5: 2D two-species five-moment + Maxwell (8 field components evolved by the PHM fluxes, Lorentz/current "
                      "sources), degree 3, 1024x512 elements"),
    "N3D": dict(dim=3, p=3, nx=[64, 64, 64], left=[0.0, -5.0, -5.0], right=[20.0, 5.0, 5.0], gamma=5.0 / 3.0, ic="two_fluid", n_species=2,
                fields=True, sources=dict(epsilon0=1.0, chi=1.0, charge_over_mass=[1.0 / 25.0, -1.0]),
                maxwell=dict(light_speed=10.0, chi=1.0, gamma=1.0),
                label="north-star shape: 3D two-species five-moment + Maxwell (8 field components evolved, Lorentz/current sources), "
                      "degree 3, 64^3 elements"),
    "N3D128": dict(dim=3, p=3, nx=[128, 128, 128], left=[0.0, -5.0, -5.0], right=[40.0, 15.0, 15.0], gamma=5.0 / 3.0, ic="two_fluid",
                   n_species=2, fields=True, sources=dict(epsilon0=1.0, chi=1.0, charge_over_mass=[1.0 / 25.0, -1.0]),
                   maxwell=dict(light_speed=10.0, chi=1.0, gamma=1.0),
                   label="north-star shape at the size SURVEY 8(d) names for one GPU: 128^3 elements, 2.4e9 DoFs, 19.3 GB per state "
                         "vector (capacity run: scripts/stage_rate.py; not a default bench line)"),
    "C2c": dict(dim=2, p=3, nx=[512, 512], left=[0.0, -5.0], right=[10.0, 5.0], gamma=1.4, ic="vortex", mapping="wavy",
                label="C2 on curved elements: the 512x512 degree-3 box pushed through a smooth periodic mapping "
                      "(general-geometry kernels, metric terms read from HBM)"),
    "C4": dict(dim=3, p=4, nx=[128, 128, 16], left=[0.0, -5.0, -5.0], right=[10.0, 5.0, -3.75], gamma=1.4, ic="vortex",
               label="C4 at its stated size when run on 8 GPUs (--gpus 8 --workload C4): 3D Euler periodic vortex, degree 4, 128^3 "
                     "elements as eight 128x128x16 slabs, 16.4 MB of face traces per direction and RHS"),
    "C4s": dict(dim=3, p=4, nx=[64, 64, 64], left=[0.0, -5.0, -5.0], right=[10.0, 5.0, 5.0], gamma=1.4, ic="vortex",
                label="3D Euler periodic vortex, degree 4, 64^3 elements per GPU (C4 shape)"),
}


OTHER_WORKLOADS = ["C2", "C3", "C4s", "C5s"]   # BASELINE configs 2-5 at their stated shapes, one GPU's share


def measured_traffic(workload):
    """DRAM bytes per stage-kernel launch from the committed ncu capture of this workload (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return int(json.load(f)[workload]["bytes"])
    except Exception:
        return None


def measured_fp64_instr(workload):
    """FP64 warp instructions of one stage-kernel launch from the committed SASS histogram of this workload, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return int(json.load(f)[workload]["fp64_warp_instr"])
    except Exception:
        return None


def fp64_figure(workload, avg_stage_ms, device):
    """The co-bound the HBM roofline does not show: FP64 warp instructions of the dominant kernel per second against the
    FMA-chain rate of this GPU measured in the same run (measure the FP64 peak before quoting a %)."""
    import ctypes as C
    from warpii_b200 import lib
    peak = C.c_double(0)
    if lib().warpii_gpu_measure_fp64_peak(device, C.byref(peak)) != 0 or peak.value <= 0:
        return None
    out = {"peak_tflops": 2.0 * peak.value / 1e12, "peak_warp_instr_per_s": peak.value / 32.0,
           "peak_source": "measured in this run: 8 independent FMA chains per thread, 8 x 256 threads per SM"}
    n = measured_fp64_instr(workload)
    if n is not None and avg_stage_ms > 0:
        rate = n / (avg_stage_ms * 1e-3)
        out.update({"fp64_warp_instr_per_launch": n, "achieved_warp_instr_per_s": rate, "frac": rate / (peak.value / 32.0)})
    return out


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Clocks and throttle reasons for the timed region.

    The SM clock DURING the region is measured on the device itself: while stage timing is on, the library enqueues a
    40 us probe kernel after every batch of steps that compares clock64 with the global timer
    (warpii_gpu_sm_clock_probes).  NVML / nvidia-smi are only queried immediately BEFORE and AFTER the region: a
    single NVML query inside it stalled the 2-GPU NCCL run by ~6 ms (20 % of a 40-step region; measured, see
    profiles/README.md), so in-region polling would make the number wrong rather than safe.  An average probe clock
    equal to the maximum clock plus empty reason sets on both sides excludes a slowdown in between.
    """
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.nvml, self.reasons, self.sm_max, self.sm_nvml = index, None, set(), None, []
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def poll(self):
        """One NVML sample (call right before / right after the timed region, GPU still busy or just released)."""
        if self.nvml:
            n = self.nvml
            try:
                self.sm_nvml.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                try:
                    r = n.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for name, bit in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            return
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm,"
                                  "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                  "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
            r = [c.strip() for c in out.strip().split(",")]
            self.sm_nvml.append(float(r[0]))
            self.sm_max = float(r[1])
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)
        except Exception:
            pass

    def result(self, probes_mhz):
        probes = [float(v) for v in probes_mhz]
        return {"sm_mhz": float(np.median(probes)) if probes else (float(np.median(self.sm_nvml)) if self.sm_nvml else None),
                "sm_max_mhz": self.sm_max, "samples": len(probes), "sm_mhz_min": min(probes) if probes else None,
                "sm_mhz_before_after_nvml": self.sm_nvml, "reasons": sorted(self.reasons),
                "source": "device clock64/globaltimer probes inside the timed region; NVML before and after"}


def base_config(w, world):
    """The part of `config` both arms print identically (the driver compares it)."""
    NN = (w["p"] + 1) ** w["dim"]
    nc = 5 * w.get("n_species", 1) + (8 if w.get("fields") else 0)
    n_local = int(np.prod(w["nx"])) * NN * nc
    label = w["label"]
    if world > 1:
        label += f" per GPU ({w['nx'][-1] * world} layers in total, slab-sharded)"
    return {"workload": label, "n_dofs": int(n_local * world), "state_bytes_per_gpu": int(8 * n_local)}


def build_ic(w, xyz):
    if w["ic"] == "vortex":
        return cases.to_state(cases.isentropic_vortex(w["gamma"])(xyz), w["gamma"])
    if w["ic"] in ("kh", "kh_steep"):
        k = 1.2 * np.pi
        if w["ic"] == "kh":
            return cases.to_state(cases.kelvin_helmholtz(k)(xyz), w["gamma"])
        x, y = xyz[..., 0], xyz[..., 1]
        steep = 4.0 * w["nx"][0] / (w["right"][0] - w["left"][0])      # transition over about one element
        prim = np.zeros(x.shape + (5,))
        prim[..., 0] = 0.65 + 0.35 * np.tanh(-(x - 0.003 * np.sin(k * y)) * steep)
        prim[..., 2] = 0.1 * np.tanh(-x * steep)
        prim[..., 4] = 1.0
        return cases.to_state(prim, w["gamma"])
    if w["ic"] == "two_fluid":
        # smooth two-fluid wave: ion (mass 25) and electron fluids, non-zero B so that the Lorentz term is active
        s = np.sin(2 * np.pi * xyz[..., 0] / 20.0) * np.cos(2 * np.pi * xyz[..., 1] / 10.0)
        u = np.zeros((xyz.shape[0], 18, xyz.shape[1]))
        for sp, (rho0, vel, p0) in enumerate([(25.0, (0.05, -0.02, 0.01), 1.0), (1.0, (-0.2, 0.1, 0.05), 1.0)]):
            prim = np.zeros(xyz.shape[:-1] + (5,))
            prim[..., 0] = rho0 * (1 + 0.1 * s)
            for d in range(3):
                prim[..., 1 + d] = vel[d] * (1 + 0.2 * s)
            prim[..., 4] = p0 * (1 + 0.05 * s)
            cases.to_state(prim, w["gamma"], nc=18, species=sp, u=u)
        for c, amp in enumerate([0.01, -0.02, 0.015, 0.1, -0.05, 0.2, 0.0, 0.0]):
            u[:, 10 + c, :] = amp * (1 + 0.3 * s)
        return u
    raise ValueError(w["ic"])


def species_kwargs(w):
    kw = dict(n_species=w.get("n_species", 1), fields_enabled=w.get("fields", False))
    if w.get("periodic") is not None:
        kw.update(periodic=w["periodic"], bc_kinds=w.get("bc"))
    return kw


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_run(w, steps, warmup, threads, budget_s, rows=None):
    """Time SSPRK2 steps of the oracle on a bounded sample: the workload's mesh cut to `rows` element rows in the
    last dimension (same h, same degree, same initial condition, periodic)."""
    import oracle
    from oracle import Oracle
    oracle.build()
    dim, nx = w["dim"], list(w["nx"])
    full_last = nx[-1]
    if rows is None:
        # calibrate on a thin slab, then size the sample so that (steps + warmup) steps fit in the budget
        probe = 4
        nxp = nx[:-1] + [probe]
        right = list(w["right"])
        right[-1] = w["left"][-1] + (w["right"][-1] - w["left"][-1]) * probe / full_last
        o = Oracle(dim, w["p"], nxp, w["left"], right, gamma=w["gamma"], threads=threads, **species_kwargs(w))
        if w.get("sources"):
            o.set_sources(True, **w["sources"])
        if w.get("maxwell"):
            o.set_maxwell(True, **w["maxwell"])
        u = build_ic(w, o.node_coords())
        dt = o.recommend_dt(u)
        t0 = time.perf_counter()
        o.ssprk2_step(u, dt, 0.0)
        per_row = (time.perf_counter() - t0) / probe
        rows = int(max(probe, min(full_last, budget_s / max(per_row * (steps + warmup), 1e-9))))
    nxs = nx[:-1] + [rows]
    right = list(w["right"])
    right[-1] = w["left"][-1] + (w["right"][-1] - w["left"][-1]) * rows / full_last
    o = Oracle(dim, w["p"], nxs, w["left"], right, gamma=w["gamma"], threads=threads, **species_kwargs(w))
    if w.get("sources"):
        o.set_sources(True, **w["sources"])
    if w.get("maxwell"):
        o.set_maxwell(True, **w["maxwell"])
    u = build_ic(w, o.node_coords())
    t = 0.0
    for _ in range(warmup):
        dt = o.recommend_dt(u)
        o.ssprk2_step(u, dt, t)
        t += dt
    t0 = time.perf_counter()
    for _ in range(steps):
        dt = o.recommend_dt(u)
        o.ssprk2_step(u, dt, t)
        t += dt
    el = time.perf_counter() - t0
    n_dofs = o.n_dofs
    return {"value": 2.0 * n_dofs * steps / el, "ms_per_step": 1e3 * el / steps, "n_dofs": int(n_dofs),
            "sample": f"{steps} SSPRK2 steps on {'x'.join(str(v) for v in nxs)} of the {'x'.join(str(v) for v in nx)} elements "
                      f"(same h, degree, IC; {n_dofs} DoFs)"}


def run_reference(args, w):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    r = cpu_run(w, args.steps, args.warmup, threads, budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": base_config(w, world),
        "details": {"note": "CPU restatement (oracle/) of the reference algorithm on all host cores of rank 0, on a bounded sample of the "
                            "workload (same h, degree, initial condition); the reference itself needs deal.II 9.5.1 + MPI and cannot be "
                            "built in this image"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def make_solver(w, world, rank, local_rank):
    """Solver, per-GPU mesh and geometry-bytes figure of one workload (weak scaling: one workload-sized slab per GPU)."""
    from warpii_b200 import BoxSolver
    dim, p, gamma = w["dim"], w["p"], w["gamma"]
    nx = list(w["nx"])
    left, right = list(w["left"]), list(w["right"])
    nx[-1] *= world
    right[-1] = left[-1] + (right[-1] - left[-1]) * world
    geo_bytes_per_dof = 0.0
    if w.get("mapping"):
        # general geometry (warpii_gpu_set_geometry): same box connectivity and patch numbering, curved support points
        if world > 1:
            raise SystemExit("general-geometry workloads run on one GPU")
        import mesh_cases as mc
        from warpii_b200 import box_tables, elems_per_block
        from warpii_b200.capi import MeshSolver, mapped_metrics
        tab = box_tables(dim, nx, [1] * dim, group=elems_per_block(dim, p))
        ref = mc.ref_nodes(dim, p)
        l2g = tab["local_to_global"]
        box_xyz = np.zeros((len(l2g), ref.shape[0], dim))
        rem = l2g.copy()
        for d in range(dim):
            box_xyz[:, :, d] = left[d] + ((rem % nx[d])[:, None] + ref[None, :, d]) * ((right[d] - left[d]) / nx[d])
            rem //= nx[d]
        mesh = {"face_neighbor": tab["face_neighbor"], "neighbor_face": None, "bf_elem": [], "bf_side": [], "bf_id": []}
        geo = mapped_metrics(dim, p, mc.wavy(left, right, 0.02)(box_xyz), mesh["face_neighbor"])
        g = MeshSolver(dim, p, mesh, geo, gamma=gamma, device=local_rank, **species_kwargs(w))
        g.node_coords = lambda: box_xyz   # the initial condition is a function of the reference-box position
        del geo
        K, NN, NF = dim * dim, (p + 1) ** dim, (p + 1) ** (dim - 1)
        # per node and stage: Ja (K) + 1/Jdet (+ the power-iteration value in the stage that fuses the CFL reduction),
        # plus the face tables of the element ((dim+1) doubles per face node)
        geo_bytes_per_dof = (8.0 * (K + 1 + 0.5) + 8.0 * 2 * dim * (dim + 1) * NF / NN) / g.nc
    else:
        g = BoxSolver(dim, p, nx, left, right, gamma=gamma, rank=rank, n_ranks=world, device=local_rank, **species_kwargs(w))
    if w.get("sources"):
        g.set_sources(True, **w["sources"])
    if w.get("maxwell"):
        g.set_maxwell(True, **w["maxwell"])
    return g, nx, geo_bytes_per_dof


def stage_kernel_name(w, g):
    from warpii_b200 import lib as _lib
    dim, np1 = w["dim"], w["p"] + 1
    if w.get("mapping"):
        return f"wgpu::stage_kernel_general<{dim},{np1}>"
    L = _lib()
    pencil = hasattr(L, "warpii_gpu_stage_kernel_is_pencil") and L.warpii_gpu_stage_kernel_is_pencil(dim, w["p"]) == 1
    return f"wgpu::{'pencil_stage_kernel' if pencil else 'stage_kernel'}<{dim},{np1}>"


N_REPEAT = 5


def device_resident(g, args, stream, barrier, max_over_ranks, steps, warmup, repeats):
    """`repeats` timed regions of exactly `steps` SSPRK2 steps each (device-resident time loop, CUDA events on the library's
    stream bracketed by barrier + synchronize, max over ranks); returns the median region and the per-launch stage time."""
    import torch
    t = 0.0
    t, _ = g.advance_to(t, 1e30, max_steps=max(warmup, 1))
    regions, stage_tot, stage_n, launches = [], 0.0, 0, 0
    for _ in range(repeats):
        barrier()
        g.stage_timing(True)
        l0 = g.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t, done = g.advance_to(t, 1e30, max_steps=steps)
        e1.record(stream)
        barrier()
        assert done == steps
        regions.append(max_over_ranks(e0.elapsed_time(e1)))
        launches = g.launch_count() - l0
        ms, n = g.stage_timing(False)
        stage_tot += ms
        stage_n += n
    return t, float(np.median(regions)), regions, stage_tot / max(stage_n, 1), stage_n, launches


def parity_check_sharded(world, rank, local_rank, dist, torch):
    """N > 1: one RHS of a small 3-D mesh sharded over all ranks (NCCL halo exchange, interface / interior launches) against
    the CPU oracle on the unsharded mesh.  Outside the timed region; the criterion is the parity tests' (dgsem_cases.py)."""
    from warpii_b200 import BoxSolver, nccl_unique_id
    import oracle
    from oracle import Oracle
    dim, p, gamma = 3, 3, 1.4
    nx = [8, 8, 4 * world]
    left, right = [0.0, -5.0, -5.0], [10.0, 5.0, 5.0]
    g = BoxSolver(dim, p, nx, left, right, gamma=gamma, rank=rank, n_ranks=world, device=local_rank)
    uid = torch.tensor(list(nccl_unique_id()) if rank == 0 else [0] * 128, dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    g.attach_comm(bytes(uid.cpu().tolist()))
    o = Oracle(dim, p, nx, left, right, gamma=gamma, threads=max(1, (os.cpu_count() or 8) // world))
    u = o.project(cases.isentropic_vortex(gamma))
    want, _ = o.rhs(u)
    g.upload(0, np.ascontiguousarray(u[g.l2g]))
    g.rhs(1, 0)
    got = g.download(1)
    g.close()
    h = [(r - l) / n for l, r, n in zip(left, right, nx)]
    scale = cases.summand_scale(u, gamma, dim, h, oracle.diff_matrix(p + 1))
    # per-rank sums of squares -> global norms
    num = np.array([np.sum((got[:, c] - want[g.l2g][:, c]) ** 2) for c in range(5)])
    t = torch.tensor(num, dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    err = np.sqrt(t.cpu().numpy())
    den = np.array([np.linalg.norm(want[:, c]) for c in range(5)])
    bound = np.maximum(1e-12 * den, cases.RHS_ULPS * 2.0 ** -52 * scale)
    live = den > 1e-9 * scale
    plain = np.where(live, err / np.where(den > 0, den, 1.0), 0.0)
    return {"mesh": f"3D p=3 {nx[0]}x{nx[1]}x{nx[2]} vortex sharded over {world} ranks, one RHS vs CPU oracle",
            "rel_l2_per_component": [float(v) for v in plain], "tolerance": 1e-12, "pass": bool((err <= bound).all()),
            "note": "component 3 (z-momentum of a flow extruded in z) vanishes identically and is judged against the differenced terms"}


def bind_to_gpu_numa_node(device_index):
    """Pin this rank's threads to the CPUs NVML lists as local to its GPU, BEFORE the pinned host buffer is allocated, so that
    the buffer's pages land on that NUMA node (first touch).  What `mpirun --bind-to` / `numactl` do for an MPI rank of the
    reference; matters for the e2e leg at N > 1, where every rank streams its whole state over its own PCIe link each step.
    Returns a short description for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[device_index]) if visible and visible.split(",")[device_index].isdigit() else device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        local = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (int(wd) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        if not cpus:
            return "GPU-local CPUs not in this process's cpuset: unbound"
        if cpus == allowed:
            return "all %d allowed CPUs are GPU-local" % len(allowed)
        os.sched_setaffinity(0, cpus)
        return "bound to %d of %d allowed CPUs (NVML CPU affinity of GPU %d)" % (len(cpus), len(allowed), index)
    except Exception as exc:   # no NVML, no affinity support: run unbound
        return "unbound (%s)" % type(exc).__name__


def run_ours(args, w):
    numa = bind_to_gpu_numa_node(int(os.environ.get("LOCAL_RANK", "0"))) if int(os.environ.get("WORLD_SIZE", "1")) > 1 else "single rank: unbound"
    import torch
    import torch.distributed as dist

    from warpii_b200 import nccl_unique_id

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    dim, p = w["dim"], w["p"]
    g, nx, geo_bytes_per_dof = make_solver(w, world, rank, local_rank)
    if world > 1:
        if rank == 0:
            uid = torch.tensor(list(nccl_unique_id()), dtype=torch.uint8, device="cuda")
        else:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        g.attach_comm(bytes(uid.cpu().tolist()))
    u0 = build_ic(w, g.node_coords())
    n_dofs_local = g.n_dofs
    n_dofs_total = n_dofs_local * world
    host = torch.empty(n_dofs_local, dtype=torch.float64).pin_memory()
    host_np = host.numpy()
    host_np[:] = u0.reshape(-1)
    del u0
    g.upload(0, host_np)
    stream = torch.cuda.ExternalStream(g.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        g.synchronize()
        torch.cuda.synchronize()

    # ---- device-resident throughput -----------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    g.advance_to(0.0, 1e30, max_steps=1)   # (every rank: the step contains the halo exchange and the all-reduce)
    if rank == 0 and not args.no_clocks:
        sampler.poll()      # under load (a warm-up step); the remaining warm-up steps and a barrier follow before the timed region
    g.upload(0, host_np)
    t, ms, regions, avg_stage_ms, stage_n, launches = device_resident(g, args, stream, barrier, max_over_ranks, args.steps,
                                                                      args.warmup, N_REPEAT)
    clocks = None
    if rank == 0 and not args.no_clocks:
        sampler.poll()
        clocks = sampler.result(g.sm_clock_probes())
    value = 2.0 * n_dofs_total * args.steps / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ------------------------------------------------
    # every step: pinned host state -> HBM, recommend_dt + one SSPRK2 step, HBM -> pinned host state
    import ctypes as C
    from warpii_b200 import lib as _lib
    L = _lib()
    hp = host_np.ctypes.data_as(C.POINTER(C.c_double))
    e2e_steps = max(3, min(args.steps, 10))

    # warpii_gpu_host_ssprk2_step: the state goes host -> HBM -> host every step; dt is the fused CFL result of the previous
    # step (= recommend_dt of the state being uploaded).  Periodic workloads overlap the two transfers and the
    # stages slab by slab; with boundary faces the call runs upload, step, download in sequence.
    def e2e_step(tt, dt_now):
        dt_next = g.host_step(host_np, host_np, dt_now, tt)
        return tt + dt_now, dt_next

    assert L.warpii_gpu_download_state(g.ctx, 0, hp, None) == 0
    dt_e2e = g.recommend_dt(0)
    for _ in range(2):
        t, dt_e2e = e2e_step(t, dt_e2e)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(e2e_steps):
        t, dt_e2e = e2e_step(t, dt_e2e)
    e1.record(stream)
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_wall = max_over_ranks((time.perf_counter() - t0) * 1e3)
    e2e_ms = max(e2e_ms, e2e_wall)   # the host-side copies are synchronous: count whichever clock saw more
    e2e_value = 2.0 * n_dofs_total * e2e_steps / (e2e_ms * 1e-3)
    assert np.isfinite(host_np).all()
    fv_frac = float((g.shock_indicator(0) > 0.0).mean())   # elements (x species) whose subcell-FV blend is active now
    kernel = stage_kernel_name(w, g)
    g.close()
    del host, host_np

    parity = parity_check_sharded(world, rank, local_rank, dist, torch) if world > 1 else None

    # ---- the other BASELINE shapes, same run, same build (N = 1 only) ------------------------------------------------
    others = {}
    if world == 1 and not args.no_other_workloads:
        peak_o, _ = measured_peaks()
        for name in OTHER_WORKLOADS:
            if name == args.workload:
                continue
            wo = WORKLOADS[name]
            go, _, _ = make_solver(wo, 1, 0, local_rank)
            uo = build_ic(wo, go.node_coords())
            go.upload(0, uo)
            del uo
            so = torch.cuda.ExternalStream(go.stream(), device=torch.device("cuda", local_rank))

            def barrier_o():
                go.synchronize()
                torch.cuda.synchronize()

            steps_o = max(3, args.steps // 2)
            _, ms_o, _, stage_o, _, _ = device_resident(go, args, so, barrier_o, lambda x: x, steps_o, 3, 3)
            others[name] = {"workload": wo["label"], "n_dofs": int(go.n_dofs), "value": 2.0 * go.n_dofs * steps_o / (ms_o * 1e-3),
                            "unit": UNIT, "ms_per_step": ms_o / steps_o, "steps": steps_o, "stage_kernel": stage_kernel_name(wo, go),
                            "stage_kernel_ms": stage_o, "roofline_frac": BYTES_PER_DOF_UPDATE * go.n_dofs / (stage_o * 1e-3) / 1e9 / peak_o}
            go.close()

    if rank == 0:
        peak, peak_src = measured_peaks()
        # dominant kernel = the fused stage kernel; algorithmic bytes per launch = 20 B (SSPRK2 average) x local DoFs
        bytes_per_update = BYTES_PER_DOF_UPDATE + geo_bytes_per_dof
        achieved = bytes_per_update * n_dofs_local / (avg_stage_ms * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": base_config(w, world),
            "details": {"l2_policy": "state (2 x %.0f MB per GPU) exceeds the 126 MB L2; no flush" % (8e-6 * n_dofs_local),
                        "parallelism": f"elements sharded over {world} GPU(s), NCCL send/recv halo + allreduce(max) dt",
                        "timed_regions": f"{N_REPEAT} x {args.steps} steps, median reported",
                        "host_numa_binding_rank0": numa,
                        "fv_blend_active_fraction_rank0": fv_frac},
            "timed_regions_ms": regions,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(8 * n_dofs_local),
                    "d2h_bytes_per_step": int(8 * n_dofs_local), "steps": e2e_steps,
                    "note": "per step: warpii_gpu_host_ssprk2_step(pinned host state in, pinned host state out): whole state host -> "
                            "HBM, SSPRK2 step with fused CFL (next dt), whole state HBM -> host; transfers and stages overlapped "
                            "slab by slab on three streams where the mesh is periodic"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload) if world == 1 else None,
                         "kernel": kernel, "peak_source": peak_src,
                         "avg_launch_ms": avg_stage_ms, "launches_timed": int(stage_n),
                         "algorithmic_bytes_per_launch": bytes_per_update * n_dofs_local,
                         "algorithmic_bytes_per_dof_update": bytes_per_update},
        }
        if w.get("maxwell") and w["dim"] == 3:
            # in 3-D the field system is a second launch over the same element range
            line["roofline"]["launch"] = ("one stage = %s + wgpu::maxwell_kernel<3,%d>, timed together; achieved, avg_launch_ms and "
                                          "traffic are per stage" % (kernel, w["p"] + 1))
        if parity is not None:
            line["parity_check"] = parity
        if others:
            line["other_workloads"] = others
        if world == 1:
            fp = fp64_figure(args.workload, avg_stage_ms, local_rank)
            if fp is not None:
                if "frac" in fp:   # the HBM-roofline fraction this kernel would reach with the FP64 pipe 100 % busy
                    t_fp64_ms = 1e3 * fp["fp64_warp_instr_per_launch"] / fp["peak_warp_instr_per_s"]
                    t_hbm_ms = 1e3 * bytes_per_update * n_dofs_local / (peak * 1e9)
                    fp["note"] = ("DFMA+DMUL+DADD+DSETP of the captured launch (profiles/*_opcodes.txt) / measured kernel time; "
                                  "with the FP64 pipe 100 %% busy the launch would take %.3f ms = %.2f of the HBM roofline"
                                  % (t_fp64_ms, t_hbm_ms / t_fp64_ms))
                    fp["hbm_frac_at_fp64_peak"] = t_hbm_ms / t_fp64_ms
                line["fp64"] = fp
        if world == 1 and not args.no_cpu_baseline and not w.get("mapping"):
            nthreads = os.cpu_count() or 1
            cb = cpu_run(w, 1, 0, nthreads, budget_s=12.0)
            cb1 = cpu_run(w, 1, 0, 1, budget_s=8.0)
            line["cpu_baseline"] = {"value": cb["value"], "unit": UNIT, "cores": nthreads, "kind": "port", "sample": cb["sample"],
                                    "single_core_value": cb1["value"], "single_core_sample": cb1["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)   # 30 N3D steps = 0.23 s per timed region
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="N3D", choices=sorted(WORKLOADS))
    ap.add_argument("--no-other-workloads", action="store_true", help="skip the other BASELINE shapes (other_workloads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="diagnostic: do not sample clocks during the timed region")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
