"""Python loader for the CPU oracle (``oracle/_build/liboracle.so``).

TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this package, and only as the
checker or as the timed CPU baseline.  PARITY UNPINNED by reference tests — see ``oracle.h``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from instaslice_b200.engine import PROFILE_DTYPE, REQUEST_DTYPE, RESULT_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
ST_VETO_REQUEUE = 100
_lib = None


def build():
    """g++ the two restatements into oracle/_build/liboracle.so (no reference sources are involved:
    the reference is Go and cannot be compiled in this image)."""
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        p = C.c_void_p
        L.orc_start_for.restype, L.orc_start_for.argtypes = C.c_uint8, [p, C.c_uint32, C.c_uint8]
        L.orc_f_new.restype, L.orc_f_new.argtypes = p, [C.c_uint32, p, C.c_uint32, p, C.c_uint32]
        L.orc_f_new_tables.restype, L.orc_f_new_tables.argtypes = p, [C.c_uint32, p, C.c_uint32, C.c_uint32, p, p, C.c_uint32]
        L.orc_fast_new_tables.restype, L.orc_fast_new_tables.argtypes = p, [C.c_uint32, p, C.c_uint32, C.c_uint32, p, p, C.c_uint32, C.c_uint32]
        L.orc_f_delete.restype, L.orc_f_delete.argtypes = None, [p]
        L.orc_f_add_prepared.restype, L.orc_f_add_prepared.argtypes = C.c_int, [p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int64]
        L.orc_f_add_allocation.restype, L.orc_f_add_allocation.argtypes = C.c_int, [p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64]
        L.orc_f_place.restype, L.orc_f_place.argtypes = C.c_int, [p, C.c_uint32, p, p, C.c_int]
        L.orc_f_occupancy.restype, L.orc_f_occupancy.argtypes = None, [p, p]
        L.orc_f_num_allocations.restype, L.orc_f_num_allocations.argtypes = C.c_uint64, [p]
        L.orc_fast_new.restype, L.orc_fast_new.argtypes = p, [C.c_uint32, p, C.c_uint32, p, C.c_uint32, C.c_uint32]
        L.orc_fast_delete.restype, L.orc_fast_delete.argtypes = None, [p]
        L.orc_fast_load.restype, L.orc_fast_load.argtypes = None, [p, p]
        L.orc_fast_place.restype, L.orc_fast_place.argtypes = C.c_int, [p, C.c_uint32, p, p]
        L.orc_fast_occupancy.restype, L.orc_fast_occupancy.argtypes = None, [p, p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def start_for(row: np.ndarray, quirks: int, occ: int) -> int:
    """One (profile row, occupancy byte) -> start in {0..7, 9}  (:343-383)."""
    row = np.ascontiguousarray(row, dtype=PROFILE_DTYPE).reshape(1)
    return int(lib().orc_start_for(_ptr(row), quirks, occ))


class Fast:
    """ref_fast.cpp: bitmask + per-profile cursor; first-fit (reference) or best-fit (extension)."""

    def __init__(self, node_off, rows, quirks=3, policy=0, node_table=None):
        """``rows`` is [n_profiles] or, for a heterogeneous cluster, [n_tables][n_profiles] with ``node_table`` [n_nodes]."""
        self.node_off = np.ascontiguousarray(node_off, dtype=np.uint32)
        self.rows = np.ascontiguousarray(rows, dtype=PROFILE_DTYPE)
        self.G = int(self.node_off[-1])
        if self.rows.ndim == 1:
            self._h = lib().orc_fast_new(len(self.node_off) - 1, _ptr(self.node_off), len(self.rows), _ptr(self.rows), quirks, policy)
        else:
            self.node_table = np.ascontiguousarray(node_table, dtype=np.uint8)
            self._h = lib().orc_fast_new_tables(len(self.node_off) - 1, _ptr(self.node_off), self.rows.shape[0], self.rows.shape[1],
                                                _ptr(self.rows), _ptr(self.node_table), quirks, policy)

    def load(self, occ):
        occ = np.ascontiguousarray(occ, dtype=np.uint8)
        assert len(occ) == self.G
        lib().orc_fast_load(self._h, _ptr(occ))

    def place(self, requests) -> np.ndarray:
        requests = np.ascontiguousarray(requests, dtype=REQUEST_DTYPE)
        out = np.zeros(len(requests), dtype=RESULT_DTYPE)
        lib().orc_fast_place(self._h, len(requests), _ptr(requests), _ptr(out))
        return out

    def occupancy(self) -> np.ndarray:
        out = np.empty(self.G, dtype=np.uint8)
        lib().orc_fast_occupancy(self._h, _ptr(out))
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_fast_delete(self._h)
            self._h = None


class Faithful:
    """ref_faithful.cpp: string-keyed CRD objects, rescans per pod — the reference as written."""

    def __init__(self, node_off, rows, quirks=3, node_table=None):
        self.node_off = np.ascontiguousarray(node_off, dtype=np.uint32)
        self.rows = np.ascontiguousarray(rows, dtype=PROFILE_DTYPE)
        self.G = int(self.node_off[-1])
        if self.rows.ndim == 1:
            self._h = lib().orc_f_new(len(self.node_off) - 1, _ptr(self.node_off), len(self.rows), _ptr(self.rows), quirks)
        else:
            self.node_table = np.ascontiguousarray(node_table, dtype=np.uint8)
            self._h = lib().orc_f_new_tables(len(self.node_off) - 1, _ptr(self.node_off), self.rows.shape[0], self.rows.shape[1],
                                             _ptr(self.rows), _ptr(self.node_table), quirks)

    def add_prepared(self, gpu, start, size, pod_id=-1):
        return lib().orc_f_add_prepared(self._h, gpu, start, size, pod_id)

    def add_allocation(self, gpu, start, size, pod_id):
        return lib().orc_f_add_allocation(self._h, gpu, start, size, pod_id)

    def load_occupancy_as_dangling(self, occ):
        """Express an occupancy byte array as dangling Prepared slices (maximal runs of busy slices)."""
        for g, b in enumerate(np.asarray(occ, dtype=np.uint8)):
            b = int(b)
            s = 0
            while s < 8:
                if (b >> s) & 1:
                    e = s
                    while e < 8 and (b >> e) & 1:
                        e += 1
                    self.add_prepared(g, s, e - s, -1)
                    s = e
                else:
                    s += 1

    def place(self, requests, all_nodes=False) -> np.ndarray:
        requests = np.ascontiguousarray(requests, dtype=REQUEST_DTYPE)
        out = np.zeros(len(requests), dtype=RESULT_DTYPE)
        lib().orc_f_place(self._h, len(requests), _ptr(requests), _ptr(out), 1 if all_nodes else 0)
        return out

    def occupancy(self) -> np.ndarray:
        out = np.empty(self.G, dtype=np.uint8)
        lib().orc_f_occupancy(self._h, _ptr(out))
        return out

    def num_allocations(self) -> int:
        return int(lib().orc_f_num_allocations(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_f_delete(self._h)
            self._h = None
