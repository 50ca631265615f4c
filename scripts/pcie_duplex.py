"""Host <-> device copy rates of this box with pinned memory: each direction alone and both at once (the ceiling of the e2e
leg of bench.py, which moves the whole state both ways every step).  usage (GPU box): python scripts/pcie_duplex.py [GB]"""
import sys
import time

import torch

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 2.4
n = int(gb * 1e9 / 8)
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
h_in.fill_(1.0)
d_a = torch.empty(n, dtype=torch.float64, device="cuda")
d_b = torch.ones(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d()
    d2h()


def chunked(k=32):
    # the same bytes in k interleaved chunks per direction (what a slab pipeline issues)
    m = n // k
    for i in range(k):
        with torch.cuda.stream(s1):
            d_a[i * m:(i + 1) * m].copy_(h_in[i * m:(i + 1) * m], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[i * m:(i + 1) * m].copy_(d_b[i * m:(i + 1) * m], non_blocking=True)


bytes_one = 8 * n
for name, fn, moved in (("H2D alone", h2d, bytes_one), ("D2H alone", d2h, bytes_one), ("both at once", both, 2 * bytes_one),
                        ("both, 32 chunks each", chunked, 2 * bytes_one)):
    t = timed(fn)
    print(f"{name:22s} {1e3 * t:8.2f} ms  {moved / t / 1e9:6.1f} GB/s", flush=True)
