"""Element-sharded run on >= 2 GPUs (NCCL send/recv halo, all-reduce(max) dt, all-reduce(sum) integrals) against the
global CPU oracle.  Skipped on a single-GPU box; the host-side partition logic is covered on CPU (gloo) in
tests/test_capi_cpu.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_halo_exchange_matches_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "rank 0: multi-GPU parity ok" in out.stdout and "rank 1: multi-GPU parity ok" in out.stdout


def test_cli_sharded_run_equals_single_gpu_run(tmp_path):
    """`warpii_gpu --gpus 2 <input>`: one process per GPU forked by the launcher, per-rank VTU pieces + a .pvtu index; the
    merged result is the single-GPU result bit for bit (same element kernels, bitwise identical traces on both sides)."""
    import numpy as np
    import torch
    import dgsem_cases as cases
    from test_input_file_cpu import read_input
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "warpii_b200", "bin", "warpii_gpu")
    text = read_input("inflow_channel_2d.inp").replace("set write_output = false", "set write_output = true")
    for name in ("one", "two"):
        (tmp_path / f"{name}.inp").write_text(text)
    r1 = subprocess.run([exe, "one.inp"], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r1.returncode == 0, r1.stdout[-2000:] + r1.stderr[-2000:]
    r2 = subprocess.run([exe, "--gpus", "2", "two.inp"], capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r2.returncode == 0, r2.stdout[-2000:] + r2.stderr[-2000:]
    assert r1.stdout.split("steps =")[1].split()[0] == r2.stdout.split("steps =")[1].split()[0]
    d1, d2 = tmp_path / "FiveMoment__one", tmp_path / "FiveMoment__two"
    names = sorted(os.listdir(d2))
    assert "solution_004.pvtu" in names and "solution_004.rank0.vtu" in names and "solution_004.rank1.vtu" in names
    index = (d2 / "solution_004.pvtu").read_text()
    assert 'Source="solution_004.rank1.vtu"' in index and 'Name="neutral_density"' in index and 'Name="pressure"' in index

    def by_position(files):
        pts, rho, en = [], [], []
        for f in files:
            v = cases.read_vtu(f)
            pts.append(v["Points"][:, :2]); rho.append(v["neutral_density"]); en.append(v["neutral_energy"])
        pts, rho, en = np.concatenate(pts), np.concatenate(rho), np.concatenate(en)
        # DG data: a point can appear in up to 4 elements; order by position, then by value, which is well defined
        order = np.lexsort((en, rho, pts[:, 0], pts[:, 1]))
        return pts[order], rho[order], en[order]
    p1, rho1, e1 = by_position([d1 / "solution_004.vtu"])
    p2, rho2, e2 = by_position([d2 / "solution_004.rank0.vtu", d2 / "solution_004.rank1.vtu"])
    assert np.array_equal(p1, p2) and np.array_equal(rho1, rho2) and np.array_equal(e1, e2)
    owners = np.concatenate([cases.read_vtu(d2 / f"solution_004.rank{r}.vtu")["owner"] for r in (0, 1)])
    assert set(owners.tolist()) == {0.0, 1.0}
    # a rank that cannot start (more ranks than element layers) is reported by the launcher before anything is forked
    r3 = subprocess.run([exe, "--gpus", "64", "two.inp"], capture_output=True, text=True, cwd=tmp_path, timeout=120)
    assert r3.returncode == 1 and "more ranks than element layers" in r3.stderr
