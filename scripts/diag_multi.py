"""Diagnostic (torchrun): where does the per-step time go on N GPUs?  fixed dt vs recommend_dt, stage time vs step time."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dgsem_cases as cases
from warpii_b200 import BoxSolver, nccl_unique_id
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
uid = torch.tensor(list(nccl_unique_id()) if rank == 0 else [0] * 128, dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0)
g = BoxSolver(2, 3, [512, 512 * world], [0.0, -5.0], [10.0, -5.0 + 10.0 * world], gamma=1.4, rank=rank, n_ranks=world, device=local)
g.attach_comm(bytes(uid.cpu().tolist()))
g.upload(0, cases.to_state(cases.isentropic_vortex(1.4)(g.node_coords()), 1.4))
def run(label, fixed, steps=20):
    t, _ = g.advance_to(0.0, 1e30, fixed_dt=fixed, max_steps=3)
    dist.barrier(); g.synchronize(); g.stage_timing(True)
    t0 = time.perf_counter()
    g.advance_to(t, 1e30, fixed_dt=fixed, max_steps=steps)
    g.synchronize(); dist.barrier()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    ms, n = g.stage_timing(False)
    if rank == 0: print(f"{label}: {wall:.3f} ms/step wall, stage avg {ms / max(n, 1):.3f} ms x {n // steps} per step", flush=True)
dt = g.recommend_dt(0)
run("fixed dt   ", 0.5 * dt)
run("adaptive dt", 0.0)
print(f"rank {rank}: interface elems {g.n_elems and lib if False else ''}") if False else None
dist.destroy_process_group()
