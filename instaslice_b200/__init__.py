"""instaslice_b200 — B200-native MIG-slot placement engine behind InstaSlice's allocator surface.

Only the placement hot path lives here: ``csrc/`` holds the sm_100a CUDA kernels and
the C ABI (``include/islplace.h`` -> ``libislplace.so``); the Python modules are the ctypes binding
(``engine``), the host-side mirror of the reference's allocator interface (``controller``), the profile
tables and the synthetic workload generators of the BASELINE configs.  There is no CPU fallback: importing
``engine`` without the built library raises.
"""
__all__ = ["engine", "controller", "tables", "workloads"]
__version__ = "0.1.0"
