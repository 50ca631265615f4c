"""CPU-side tests of the product: the C-ABI library loads and exports what include/*.h declares, and the host
layer (time loop, mesh tables, partitioning) behaves like the reference.  No compute call needs a GPU here."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import warpii_b200
from warpii_b200 import box_tables, host_advance

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(warpii_[a-z0-9_]+)\s*\(", text)) - {"warpii_callback_fn", "warpii_step_fn",
                                                                            "warpii_dt_fn", "warpii_cb_index_fn"})


@pytest.mark.parametrize("header", ["warpii_gpu.h", "warpii_host.h"])
def test_library_exports_every_declared_symbol(header):
    L = warpii_b200.lib()
    names = declared_functions(header)
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/{header} but not exported by libwarpii_b200.so"
    assert L.warpii_gpu_abi_version() == 1


def test_library_is_built_for_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", warpii_b200.lib_path()], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the operator must refuse to exist, not silently compute on the host."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(warpii_b200.WarpiiGpuError) as e:
        warpii_b200.BoxSolver(1, 2, [4], [0.0], [1.0])
    assert "CUDA" in str(e.value) or "cuda" in str(e.value)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under warpii_b200/ may reference it."""
    pkg = os.path.join(ROOT, "warpii_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath or os.sep + "lib" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "dgsem_oracle" not in text \
                    and "liboracle" not in text, f"{f} references the oracle"


# ---- test/timestepper_test.cc against the product's advance() --------------------------------------------
def test_host_advance_no_callbacks():
    host_advance(lambda t, dt: True, 19.0, lambda: 0.024, [])


def test_host_advance_stops_at_callbacks():
    wt, dg, pp = [], [], []
    cbs = [(0.3, wt.append, True, True), (0.1, dg.append, True, True), (0.25, pp.append, False, True)]
    host_advance(lambda t, dt: True, 1.2, lambda: 0.024, cbs)
    assert len(wt) == 5 and wt[0] == 0.0 and abs(wt[3] - 0.9) < 1e-12 and abs(wt[4] - 1.2) < 1e-12
    assert len(dg) == 13 and dg[0] == 0.0 and abs(dg[6] - 0.6) < 1e-12 and abs(dg[12] - 1.2) < 1e-12
    assert len(pp) == 5 and abs(pp[0] - 0.25) < 1e-12 and abs(pp[2] - 0.75) < 1e-12 and abs(pp[4] - 1.2) < 1e-12


def test_host_advance_no_wasted_steps():
    count = [0]

    def step(t, dt):
        count[0] += 1
        return True

    w, d = [], []
    host_advance(step, 1.2, lambda: 0.024, [(0.3, w.append, True, True), (0.3, d.append, True, True)])
    two = count[0]
    count[0] = 0
    host_advance(step, 1.2, lambda: 0.024, [(0.3, w.append, True, True)])
    assert two == 52 and count[0] == 52


def test_host_advance_matches_oracle_on_irregular_dt():
    import oracle
    seq = [0.013, 0.0071, 0.02, 0.0033]

    def run(adv):
        i = [0]
        log = []

        def dt():
            i[0] += 1
            return seq[i[0] % len(seq)]

        adv(lambda t, d: (log.append((t, d)) or True), 0.5, dt, [(0.11, lambda t: log.append(("cb", t)), True, False)])
        return log

    assert run(host_advance) == run(oracle.advance)


# ---- mesh tables ---------------------------------------------------------------------------------------------
def ref_neighbor(dim, nx, periodic, g, f):
    idx = []
    t = g
    for d in range(dim):
        idx.append(t % nx[d])
        t //= nx[d]
    d, side = f // 2, f % 2
    i = idx[d] + (1 if side else -1)
    if i < 0 or i >= nx[d]:
        if not periodic[d]:
            return -1 - f
        i %= nx[d]
    idx[d] = i
    out = 0
    for d in reversed(range(dim)):
        out = out * nx[d] + idx[d]
    return out


@pytest.mark.parametrize("dim,nx,periodic", [(1, [7], [1]), (1, [5], [0]), (2, [4, 3], [1, 0]), (2, [3, 3], [1, 1]),
                                             (3, [3, 2, 4], [0, 1, 1]), (3, [2, 2, 2], [0, 0, 0])])
def test_box_tables_single_rank(dim, nx, periodic):
    t = box_tables(dim, nx, periodic)
    n = int(np.prod(nx))
    assert t["n_local"] == n and t["n_ghost"] == 0 and t["n_interface"] == 0
    assert np.array_equal(t["local_to_global"], np.arange(n))
    nb_faces = 0
    for e in range(n):
        for f in range(2 * dim):
            v = t["face_neighbor"][e, f]
            want = ref_neighbor(dim, nx, periodic, e, f)
            if want >= 0:
                assert v == want
            else:
                b = -1 - v
                assert t["bf_elem"][b] == e and t["bf_side"][b] == f
                assert t["bf_id"][b] == f          # colorize=true: boundary id 2d+side (grid_descriptions.cc:55-56)
                nb_faces += 1
    assert nb_faces == t["n_bfaces"]


@pytest.mark.parametrize("dim,nx,periodic,R", [(1, [8], [1], 2), (2, [4, 6], [1, 1], 3), (2, [3, 4], [1, 0], 2),
                                               (3, [3, 3, 8], [1, 1, 1], 4), (3, [2, 3, 2], [1, 1, 1], 2),
                                               (3, [4, 4, 8], [1, 1, 1], 8)])
@pytest.mark.parametrize("group", [1, 4])
def test_box_tables_partition_is_consistent(dim, nx, periodic, R, group):
    """Every element owned once; ghost slots of rank a pair up with the send list of the peer, in the same order."""
    tabs = [box_tables(dim, nx, periodic, r, R, group=group) for r in range(R)]
    n = int(np.prod(nx))
    owned = np.concatenate([t["local_to_global"] for t in tabs])
    assert sorted(owned.tolist()) == list(range(n))
    for r, t in enumerate(tabs):
        l2g = t["local_to_global"]
        g2l = {int(g): i for i, g in enumerate(l2g)}
        # interface elements come first and are exactly the ones with a ghost face
        has_ghost = (t["face_neighbor"] >= t["n_local"]).any(axis=1)
        assert has_ghost[:t["n_interface"]].all() and not has_ghost[t["n_interface"]:].any()
        for e in range(t["n_local"]):
            for f in range(2 * dim):
                v = int(t["face_neighbor"][e, f])
                want = ref_neighbor(dim, nx, periodic, int(l2g[e]), f)
                if want < 0:
                    assert v < 0 and t["bf_id"][-1 - v] == f
                elif v < t["n_local"]:
                    assert int(l2g[v]) == want
                else:
                    slot = v - t["n_local"]
                    assert int(t["ghost_global_elem"][slot]) == want and int(t["ghost_side"][slot]) == (f ^ 1)
        for pi, peer in enumerate(t["peer_rank"]):
            tp = tabs[peer]
            qi = list(tp["peer_rank"]).index(r)
            mine = slice(int(t["recv_offset"][pi]), int(t["recv_offset"][pi + 1]))
            theirs = slice(int(tp["send_offset"][qi]), int(tp["send_offset"][qi + 1]))
            sent_g = tp["local_to_global"][tp["send_elem"][theirs]]
            assert np.array_equal(sent_g, t["ghost_global_elem"][mine])
            assert np.array_equal(tp["send_side"][theirs], t["ghost_side"][mine])
        del g2l


def test_patch_numbering_groups_compact_patches():
    """With group = 16 in 2D the first 16 device elements form a 4x4 patch, so 24 of their 64 faces are internal."""
    t = box_tables(2, [8, 12], [1, 1], group=16)
    l2g = t["local_to_global"]
    assert sorted(l2g.tolist()) == list(range(96))
    first = l2g[:16]
    assert sorted((first % 8).tolist()) == sorted([0, 1, 2, 3] * 4) and sorted((first // 8).tolist()) == sorted([0, 1, 2, 3] * 4)
    nb = t["face_neighbor"][:16]
    assert int(((nb >= 0) & (nb < 16)).sum()) == 2 * 24
    # a mesh that the patch shape does not divide is still a valid permutation
    t = box_tables(3, [3, 5, 2], [1, 0, 1], group=8)
    assert sorted(t["local_to_global"].tolist()) == list(range(30))


# ---- N > 1 path on CPU: gloo, world size 2 ---------------------------------------------------------------------
GLOO_WORKER = r"""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from warpii_b200 import box_tables

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, R = dist.get_rank(), 2
dim, nx, periodic, NF5 = 2, [4, 6], [1, 1], 5 * 4
t = box_tables(dim, nx, periodic, rank, R)
# a fake "trace" that encodes (global element, side): what the pack kernel would put on the wire
def trace(g, side):
    return torch.full((NF5,), float(g * 10 + side), dtype=torch.float64)
send = torch.stack([trace(int(t["local_to_global"][e]), int(s)) for e, s in zip(t["send_elem"], t["send_side"])])
ghost = torch.zeros((t["n_ghost"], NF5), dtype=torch.float64)
reqs = []
for pi, peer in enumerate(t["peer_rank"]):
    s0, s1 = int(t["send_offset"][pi]), int(t["send_offset"][pi + 1])
    r0, r1 = int(t["recv_offset"][pi]), int(t["recv_offset"][pi + 1])
    reqs.append(dist.isend(send[s0:s1].contiguous(), int(peer)))
    reqs.append(dist.irecv(ghost[r0:r1], int(peer)))
for r in reqs:
    r.wait()
want = torch.tensor([float(g * 10 + s) for g, s in zip(t["ghost_global_elem"], t["ghost_side"])], dtype=torch.float64)
assert torch.equal(ghost[:, 0], want), (ghost[:, 0], want)
# dt reduction: MAX over ranks of the local transport speed (replaces Utilities::MPI::max)
v = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(v, op=dist.ReduceOp.MAX)
assert v.item() == 2.0
dist.barrier()
dist.destroy_process_group()
print("OK", rank)
"""


def test_halo_tables_over_gloo_world_size_2(tmp_path):
    """Exchange fake face traces between two CPU ranks with the product's send/recv tables."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = str(s.getsockname()[1])
    s.close()
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                              text=True) for r in range(2)]
    for r, p in enumerate(procs):
        try:
            out, err = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            raise
        assert p.returncode == 0, err[-2000:]
        assert f"OK {r}" in out
