"""ref_py — pure-Python restatement of the reference allocator on CRD-shaped dicts.

TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/oracle.h).  PARITY UNPINNED by the
reference's own tests (none reaches the allocator); this module is one of
three independent restatements that must agree, and it is the one whose objects look like
the Instaslice custom resource (``config/crd/bases/inference.codeflare.dev_instaslices.yaml``),
so that parity tests can be written against realistic CR states.

Follows ``internal/controller/instaslice_controller.go`` of the reference (commit b34e86d):
``getStartIndexFromPreparedState`` :303-384, ``findDeviceForASlice`` :240-262,
``extractGpuProfile`` :283-300, ``FirstFitPolicy.SetAllocationDetails`` :436-453,
``extractProfileName`` :265-280 and the node loop of ``Reconcile`` :188-232.

An Instaslice is a dict::

    {"metadata": {"name": node},
     "spec": {"MigGPUUUID": {uuid: model}, "allocations": {podUID: {...}},
              "prepared": {migUUID: {...}}, "migplacement": [{"profile", "placements": [{"size","start"}],
              "giprofileid", "ciProfileid", "ciengprofileid"}]}}

with the JSON field names of ``api/v1alpha1/instaslice_types.go:23-72``.
"""
from __future__ import annotations

import re

NOT_VALID_INDEX = 9  # :248, :343
QUIRK_STRICT_BOUND = 1
QUIRK_POW2_ONLY = 2
REF_EXACT = QUIRK_STRICT_BOUND | QUIRK_POW2_ONLY
FIXED = 0


class RefPanic(Exception):
    """The Go code would panic here (index out of range)."""


def extract_profile_name(limits: dict) -> str:
    """:265-280 — last key containing "nvidia" that matches ``(\\d+g\\.\\d+gb)`` wins (map order: sorted here)."""
    name = ""
    for k in sorted(limits):
        if "nvidia" in k:
            m = re.search(r"(\d+g\.\d+gb)", k)
            if m:
                name = m.group(1)
    return name


def get_start_index_from_prepared_state(instaslice: dict, gpu_uuid: str, profile_name: str, quirks: int = REF_EXACT) -> int:
    """:303-384"""
    spec = instaslice["spec"]
    busy = [0] * 8                                                       # :306-310
    for item in spec.get("prepared", {}).values():                      # :312-320
        if item["parent"] == gpu_uuid and item.get("podUUID", "") == "":
            for i in range(int(item["size"])):
                if item["start"] + i >= 8:
                    raise RefPanic("prepared span beyond slot 7")
                busy[item["start"] + i] = 1
    for item in spec.get("allocations", {}).values():                   # :322-328, any allocationStatus
        if item["gpuUUID"] == gpu_uuid:
            for i in range(int(item["size"])):
                if item["start"] + i >= 8:
                    raise RefPanic("allocation span beyond slot 7")
                busy[item["start"] + i] = 1
    needed = 0
    starts = []
    for row in spec.get("migplacement", []):                            # :332-340, first row with that name
        if row["profile"] == profile_name:
            if not row.get("placements"):
                raise RefPanic("Placements[0] on empty list")
            needed = row["placements"][0]["size"]
            starts = [p["start"] for p in row["placements"]]
            break
    strict = bool(quirks & QUIRK_STRICT_BOUND)
    pow2 = bool(quirks & QUIRK_POW2_ONLY)
    new_start = NOT_VALID_INDEX                                          # :343
    for v in starts:                                                     # :344-381
        if v >= 8 or v < 0:
            raise RefPanic("start outside [0,8)")
        if busy[v] != 0:
            continue
        if needed == 1:
            new_start = v
            break
        handled = needed in (2, 4, 8) if pow2 else 2 <= needed <= 8
        if not handled:
            continue
        inside = (v + needed < 8) if strict else (v + needed <= 8)       # :351, :360, :370
        if not inside:
            continue
        if any(busy[v + i] for i in range(needed)):
            continue
        new_start = v
        if needed == 8 and strict:
            continue                                                     # :368-378 lacks a break (dead code under Q1)
        break
    return new_start


def extract_gpu_profile(instaslice: dict, profile_name: str):
    """:283-300 — LAST matching row wins; size of its first placement."""
    size = gi = ci = cieng = 0
    for row in instaslice["spec"].get("migplacement", []):
        if row["profile"] == profile_name:
            for p in row.get("placements", []):
                size = p["size"]
                gi, ci, cieng = row["giprofileid"], row["ciProfileid"], row["ciengprofileid"]
                break
    return size, gi, ci, cieng


def set_allocation_details(profile_name, new_start, size, pod_uuid, nodename, processed, gi, ci, cieng, namespace, pod_name, gpu_uuid):
    """FirstFitPolicy.SetAllocationDetails :436-453 (JSON field names of AllocationDetails)."""
    return {"profile": profile_name, "start": new_start, "size": size, "podUUID": pod_uuid, "gpuUUID": gpu_uuid,
            "nodename": nodename, "allocationStatus": processed, "giprofileid": gi, "ciProfileid": ci,
            "ciengprofileid": cieng, "namespace": namespace, "podName": pod_name}


def find_device_for_a_slice(instaslice: dict, profile_name: str, pod: dict, quirks: int = REF_EXACT):
    """:240-262.  Returns AllocationDetails or None ("failed to find allocatable gpu")."""
    for gpu_uuid in sorted(instaslice["spec"].get("MigGPUUUID", {})):   # :242, canonical = ascending UUID
        instaslice["spec"].setdefault("allocations", {})                 # :243-245
        new_start = get_start_index_from_prepared_state(instaslice, gpu_uuid, profile_name, quirks)
        if new_start == NOT_VALID_INDEX:
            continue
        size, gi, ci, cieng = extract_gpu_profile(instaslice, profile_name)
        return set_allocation_details(profile_name, new_start, size, pod["uid"], instaslice["metadata"]["name"], "creating",
                                      gi, ci, cieng, pod.get("namespace", "default"), pod["name"], gpu_uuid)
    return None


def reconcile_gated_pod(instaslices: list, pod: dict, profile_name: str, quirks: int = REF_EXACT, all_nodes: bool = False):
    """Node loop of Reconcile :188-232 for one gated pod that has no allocation yet.

    Returns ``("placed", [AllocationDetails...])``, ``("veto", [])`` (:198-203, RequeueAfter 1 s) or
    ``("none", [])`` (:229-232, RequeueAfter 2 s).  ``all_nodes=True`` is the literal reference behaviour
    (no break after the first successful node, Q5); the canonical semantics stops at the first node.
    """
    placed = []
    for instaslice in instaslices:                                       # :190
        alloc = find_device_for_a_slice(instaslice, profile_name, pod, quirks)
        if alloc is None:
            continue                                                     # :193-196
        for item in instaslice["spec"].get("prepared", {}).values():    # :198-203
            if item["parent"] == alloc["gpuUUID"] and item["size"] == alloc["size"] and item["start"] == alloc["start"]:
                return ("veto", placed)
        instaslice["spec"].setdefault("allocations", {})[pod["uid"]] = alloc   # :215-219
        placed.append(alloc)
        if not all_nodes:
            break
    return ("placed", placed) if placed else ("none", [])
