"""MIG profile tables as the node daemonset publishes them in ``Instaslice.Spec.Migplacement``.

The reference discovers these rows from NVML (``internal/controller/instaslice_daemonset.go:588-664``:
one ``Mig{Placements[{Size,Start}], Profile, Giprofileid, CIProfileID, CIEngProfileID}`` per supported
GI profile, CI profile id := GI profile id, CI engine profile 0).  The values below are the well-known
NVML placement tables; profile ids are ``NVML_GPU_INSTANCE_PROFILE_*`` from nvml.h.
``size`` is in memory slices, ``starts`` are the legal first slices in NVML order.
"""
from __future__ import annotations

# name, size (memory slices), legal starts in NVML order, GI profile id
A100_40GB = [
    ("1g.5gb", 1, [0, 1, 2, 3, 4, 5, 6], 0),
    ("2g.10gb", 2, [0, 2, 4], 1),
    ("3g.20gb", 4, [0, 4], 2),
    ("4g.20gb", 4, [0], 3),
    ("7g.40gb", 8, [0], 4),
    ("1g.10gb", 2, [0, 2, 4, 6], 9),
]

# H100-80GB / A100-80GB class ("1g.10gb, 7 GI slots" of BASELINE configs 2-4)
H100_80GB = [
    ("1g.10gb", 1, [0, 1, 2, 3, 4, 5, 6], 0),
    ("1g.20gb", 2, [0, 2, 4, 6], 9),
    ("2g.20gb", 2, [0, 2, 4], 1),
    ("3g.40gb", 4, [0, 4], 2),
    ("4g.40gb", 4, [0], 3),
    ("7g.80gb", 8, [0], 4),
]

# A30-24GB: four memory slices.  The reference's search hard-codes 8 slots (instaslice_controller.go:306) and the strict
# `start + size < 8` bound (Q1), so a 4-slice GPU needs nothing special: its rows simply never name a start above 3
# and `4g.24gb` (size 4, start 0) is placeable.  SURVEY 8f-3 "generalised slot width".
A30_24GB = [
    ("1g.6gb", 1, [0, 1, 2, 3], 0),
    ("2g.12gb", 2, [0, 2], 1),
    ("4g.24gb", 4, [0], 3),
]

# B200-180GB: same placement geometry as the 80GB class, names as in NVIDIA's MIG user guide (not verifiable offline)
B200_180GB = [
    ("1g.23gb", 1, [0, 1, 2, 3, 4, 5, 6], 0),
    ("1g.45gb", 2, [0, 2, 4, 6], 9),
    ("2g.45gb", 2, [0, 2, 4], 1),
    ("3g.90gb", 4, [0, 4], 2),
    ("4g.90gb", 4, [0], 3),
    ("7g.180gb", 8, [0], 4),
]

TABLES = {"a100-40gb": A100_40GB, "h100-80gb": H100_80GB, "a30-24gb": A30_24GB, "b200-180gb": B200_180GB}


def migplacement(table):
    """The table as the JSON rows of ``spec.migplacement`` (field names of instaslice_types.go:23-34)."""
    return [
        {"profile": name, "placements": [{"size": size, "start": s} for s in starts],
         "giprofileid": gi, "ciProfileid": gi, "ciengprofileid": 0}
        for name, size, starts, gi in table
    ]


def profile_index(table, name):
    """Row index the start search uses for ``name``: the FIRST row with that name (:332-340)."""
    for i, row in enumerate(table):
        if row[0] == name:
            return i
    return 0xFF
