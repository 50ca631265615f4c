"""warpii_b200: B200-native ES-DGSEM right-hand side + SSPRK2 for WarpII's FiveMoment/Euler hot path.

The product is libwarpii_b200.so (hand-written sm_100a CUDA kernels behind the C ABI of include/warpii_gpu.h,
plus the C++ host layer of warpii_b200/host/).  This Python package is only a ctypes binding used by the tests
and bench.py; there is no Python or CPU implementation of the operator behind it, and loading fails loudly if
the shared library has not been built (`python -c "import __graft_entry__ as g; g.build()"` or
`make -C warpii_b200`).
"""
from .capi import (BC_INFLOW, BC_OUTFLOW, BC_SUBSONIC_OUTFLOW, BC_WALL, FUSE_CFL, App, BoxSolver, WarpiiGpuError, box_tables, elems_per_block,
                   host_advance, lib, lib_path, nccl_unique_id)

__all__ = ["App", "BoxSolver", "WarpiiGpuError", "lib", "lib_path", "box_tables", "host_advance", "nccl_unique_id", "elems_per_block",
           "BC_WALL", "BC_OUTFLOW", "BC_INFLOW", "BC_SUBSONIC_OUTFLOW", "FUSE_CFL"]
