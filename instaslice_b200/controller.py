"""Host-side mirror of the reference's allocator interface, backed by the CUDA engine.

The reference's controller is Go and Go cannot be compiled in this image, so this module plays the part of
the Go shim described in INTEGRATION.md: same names, argument meaning and error behaviour as
``internal/controller/instaslice_controller.go`` for the allocator path, on CR-shaped dicts (JSON field names of
``api/v1alpha1/instaslice_types.go``).  Nothing here decides a placement: occupancy bytes and profile rows are
*derived* from the custom resources exactly the way the reference reads them, and every decision comes back
from ``libislplace.so`` through the C ABI.

  InstasliceReconciler.findDeviceForASlice            :240-262
  InstasliceReconciler.getStartIndexFromPreparedState :303-384   (occupancy build :306-328 is ``occupancy_byte``)
  InstasliceReconciler.extractGpuProfile              :283-300
  InstasliceReconciler.extractProfileName             :265-280
  FirstFitPolicy.SetAllocationDetails                 :436-453
  InstasliceReconciler.reconcile_gated_pod            the node loop + Prepared veto of Reconcile :188-232
  InstasliceReconciler.place_pending_pods             the batched entry the engine was built for
"""
from __future__ import annotations

import re

import numpy as np

from . import engine as E

NOT_VALID_INDEX = 9   # :248, :343
ERR_NO_GPU = "failed to find allocatable gpu"   # :261


class AllocationError(Exception):
    """The Go ``error`` value of findDeviceForASlice."""


class FirstFitPolicy:
    """:436-453 — packs the twelve arguments into AllocationDetails; it chooses nothing."""

    def SetAllocationDetails(self, profileName, newStart, size, podUUID, nodename, processed, discoveredGiprofile,
                             Ciprofileid, Ciengprofileid, namespace, podName, gpuUuid):
        return {"profile": profileName, "start": int(newStart), "size": int(size), "podUUID": podUUID, "gpuUUID": gpuUuid,
                "nodename": nodename, "allocationStatus": processed, "giprofileid": discoveredGiprofile,
                "ciProfileid": Ciprofileid, "ciengprofileid": Ciengprofileid, "namespace": namespace, "podName": podName}


class LeftToRightPolicy:
    """:456-461 — a stub in the reference: returns an empty AllocationDetails."""

    def SetAllocationDetails(self, *args):
        return {}


class RightToLeftPolicy(LeftToRightPolicy):
    """:464-469 — same stub."""


def occupancy_byte(instaslice: dict, gpu_uuid: str) -> int:
    """:306-328 — dangling Prepared (PodUUID == "") and every Allocations entry (any status) mark their slices."""
    spec = instaslice["spec"]
    busy = 0
    for item in spec.get("prepared", {}).values():
        if item["parent"] == gpu_uuid and item.get("podUUID", "") == "":
            if item["start"] + item["size"] > 8:
                raise ValueError("prepared span beyond slice 7 (the reference would panic, :316)")
            busy |= ((1 << int(item["size"])) - 1) << int(item["start"])
    for item in spec.get("allocations", {}).values():
        if item["gpuUUID"] == gpu_uuid:
            if item["start"] + item["size"] > 8:
                raise ValueError("allocation span beyond slice 7 (the reference would panic, :325)")
            busy |= ((1 << int(item["size"])) - 1) << int(item["start"])
    return busy & 0xFF


def profile_rows(migplacement: list):
    """``spec.migplacement`` -> (isl_profile records, {name: row index}).

    The start search uses the FIRST row with a given name (:332-340, ``break``); row order is kept so that
    index == position of that first row.  Duplicate starts inside a row are dropped (they cannot change the
    first hit).  A row with an empty placement list makes the reference panic (:334): rejected here.
    """
    names, table = {}, []
    for row in migplacement:
        if row["profile"] in names:
            continue
        if not row.get("placements"):
            raise ValueError("Migplacement row %r has no placements (reference panics at :334)" % row["profile"])
        names[row["profile"]] = len(table)
        table.append((row["profile"], row["placements"][0]["size"], [p["start"] for p in row["placements"]], row["giprofileid"]))
    if len(table) > E.MAX_PROFILES:
        raise ValueError("more than %d distinct profiles" % E.MAX_PROFILES)
    return E.make_profiles(table), names


class InstasliceReconciler:
    """Allocator half of the reference's ``InstasliceReconciler`` over a list of Instaslice objects.

    The engine mirrors the listed custom resources (``sync``); GPUs are in canonical order: nodes in list order,
    GPUs by ascending UUID inside a node (the reference's orders are random).
    """

    def __init__(self, instaslices: list, quirks: int = E.QUIRKS_REF_EXACT, max_batch: int = 65536, engine: E.Engine | None = None):
        self.quirks = quirks
        self.items = instaslices
        self._engine = engine
        self._max_batch = max_batch
        self.sync()

    # -- CR -> engine ---------------------------------------------------------------------------
    def sync(self):
        """Rebuild the flat inventory from the custom resources (the CR is the checkpoint)."""
        if not self.items:
            raise ValueError("no Instaslice objects")
        # every node publishes its own Migplacement (instaslice_daemonset.go:588-664): group identical tables
        self._tables, self.node_table = [], []
        for it in self.items:
            mig = it["spec"].get("migplacement", [])
            if mig not in self._tables:
                if len(self._tables) >= 8:
                    raise ValueError("more than 8 distinct per-node profile tables")
                self._tables.append(mig)
            self.node_table.append(self._tables.index(mig))
        per_table = [profile_rows(mig) for mig in self._tables]
        self.profile_names = {}
        for _rows, names in per_table:                       # profile NAME index = order of first appearance over the tables
            for name in names:
                self.profile_names.setdefault(name, len(self.profile_names))
        if len(self.profile_names) > E.MAX_PROFILES:
            raise ValueError("more than %d distinct profile names" % E.MAX_PROFILES)
        self.rows = np.zeros((len(self._tables), len(self.profile_names)), dtype=E.PROFILE_DTYPE)
        for t, (rows, names) in enumerate(per_table):
            for name, idx in names.items():
                self.rows[t, self.profile_names[name]] = rows[idx]
        self.gpu_uuid, node_off, occ = [], [0], []
        self.node_of_uuid = {}
        for n, it in enumerate(self.items):
            for uuid in sorted(it["spec"].get("MigGPUUUID", {})):
                self.gpu_uuid.append(uuid)
                self.node_of_uuid[uuid] = n
                occ.append(occupancy_byte(it, uuid))
            node_off.append(len(self.gpu_uuid))
        self.node_off = np.asarray(node_off, dtype=np.uint32)
        self.gpu_index = {u: i for i, u in enumerate(self.gpu_uuid)}
        if self._engine is None:
            self._engine = E.Engine(max_gpus=max(4096, len(self.gpu_uuid)), max_batch=self._max_batch, quirks=self.quirks)
        self._engine.load_profile_tables(self.rows)
        self._engine.load_inventory(self.node_off, np.asarray(occ, dtype=np.uint8))
        self._engine.set_node_tables(np.asarray(self.node_table, dtype=np.uint8))
        # spans of realised slices whose Allocations entry is gone: only these can trigger the veto (:198-203)
        self._has_orphans = any(
            p.get("podUUID", "") != "" and p["podUUID"] not in it["spec"].get("allocations", {})
            for it in self.items for p in it["spec"].get("prepared", {}).values())

    def update_node(self, instaslice: dict):
        """Incremental sync after ONE Instaslice object changed (an Allocations / Prepared entry appeared, changed or was
        deleted by the daemonset): recompute the occupancy bytes of that node's GPUs only and overwrite them in the engine.
        Falls back to a full ``sync`` when the node's GPU set or profile table changed."""
        n = next((i for i, it in enumerate(self.items) if it["metadata"]["name"] == instaslice["metadata"]["name"]), None)
        if n is None:
            self.items.append(instaslice)
            return self.sync()
        lo, hi = int(self.node_off[n]), int(self.node_off[n + 1])
        uuids = sorted(instaslice["spec"].get("MigGPUUUID", {}))
        if uuids != self.gpu_uuid[lo:hi] or instaslice["spec"].get("migplacement", []) != self._tables[self.node_table[n]]:
            self.items[n] = instaslice
            return self.sync()
        self.items[n] = instaslice
        self._engine.write_occupancy(lo, np.array([occupancy_byte(instaslice, u) for u in uuids], dtype=np.uint8))
        self._has_orphans = any(
            p.get("podUUID", "") != "" and p["podUUID"] not in it["spec"].get("allocations", {})
            for it in self.items for p in it["spec"].get("prepared", {}).values())

    @property
    def engine(self) -> E.Engine:
        return self._engine

    # -- reference-named helpers ----------------------------------------------------------------
    @staticmethod
    def extractProfileName(limits: dict) -> str:
        """:265-280"""
        name = ""
        for k in sorted(limits):
            if "nvidia" in k:
                m = re.search(r"(\d+g\.\d+gb)", k)
                if m:
                    name = m.group(1)
        return name

    @staticmethod
    def extractGpuProfile(instaslice: dict, profileName: str):
        """:283-300 — the LAST matching row wins; size of its first placement."""
        size = gi = ci = cieng = 0
        for row in instaslice["spec"].get("migplacement", []):
            if row["profile"] == profileName:
                for p in row.get("placements", []):
                    size, gi, ci, cieng = p["size"], row["giprofileid"], row["ciProfileid"], row["ciengprofileid"]
                    break
        return size, gi, ci, cieng

    def getStartIndexFromPreparedState(self, instaslice: dict, gpuUUID: str, profileName: str) -> int:
        """:303-384 — the occupancy byte comes from the CR, the search from the device table."""
        row = self.profile_names.get(profileName)
        if row is None:
            return NOT_VALID_INDEX
        occ = np.array([occupancy_byte(instaslice, gpuUUID)], dtype=np.uint8)
        n = next(i for i, it in enumerate(self.items) if it is instaslice or it["metadata"]["name"] == instaslice["metadata"]["name"])
        return int(self._engine.eval_starts(row | (self.node_table[n] << 8), occ)[0])       # the node's own table

    def findDeviceForASlice(self, instaslice: dict, profileName: str, policy, pod: dict) -> dict:
        """:240-262 — first GPU of ONE node with a valid start; raises AllocationError(:261) when none.

        Like the reference this does not write the allocation into the CR (:257 is commented out there); the engine's
        occupancy is left untouched as well (the tentative commit is released again).
        """
        n = next(i for i, it in enumerate(self.items) if it is instaslice)
        lo, hi = int(self.node_off[n]), int(self.node_off[n + 1])
        res = self._place([profileName], lo, hi)[0]
        if res["status"] != E.ST_PLACED:
            raise AllocationError(ERR_NO_GPU)
        self._release(res)
        return self._details(instaslice, profileName, policy, pod, res)

    # -- Reconcile's node loop, one pod (:188-232) ------------------------------------------------
    def reconcile_gated_pod(self, pod: dict, profileName: str, policy=None):
        """Returns ("placed", AllocationDetails) | ("veto", None) | ("none", None); on "placed" the allocation is
        written to the owning Instaslice (``r.Update``, :218-219).  Canonical semantics: the first node with
        capacity wins (the reference has no ``break`` there)."""
        policy = policy or FirstFitPolicy()
        res = self._place([profileName], 0, len(self.gpu_uuid))[0]
        if res["status"] != E.ST_PLACED:
            return ("none", None)
        return self._commit_or_veto(pod, profileName, policy, res)

    # -- the batched entry ------------------------------------------------------------------------
    def place_pending_pods(self, pods: list, policy=None):
        """Resolve many gated pods ``[{"uid","name","namespace","profile"}]`` in order with ONE engine call.

        Returns a list of ("placed", AllocationDetails) | ("veto", None) | ("none", None).  When the cluster holds
        realised slices whose allocation is already gone (the only state in which the reference's exact-match veto
        can fire) the pods are resolved one engine call each, so that a vetoed pod leaves no trace before the next
        one is looked at — exactly the reference's sequence.
        """
        policy = policy or FirstFitPolicy()
        if self._has_orphans:
            return [self.reconcile_gated_pod(p, p["profile"], policy) for p in pods]
        out = []
        results = self._place([p["profile"] for p in pods], 0, len(self.gpu_uuid))
        for pod, res in zip(pods, results):
            if res["status"] != E.ST_PLACED:
                out.append(("none", None))
            else:
                out.append(self._commit_or_veto(pod, pod["profile"], policy, res))
        return out

    def release(self, pod_uid: str):
        """The daemonset deleted ``Allocations[podUID]`` (instaslice_daemonset.go:261-263): free its span."""
        for n, it in enumerate(self.items):
            a = it["spec"].get("allocations", {}).pop(pod_uid, None)
            if a is not None:
                # rebuild the node's bytes from the CR (OR over every remaining entry, :306-328) instead of clearing the span blindly:
                # slices another entry still covers stay busy, exactly what the reference's next rebuild would say
                lo, hi = int(self.node_off[n]), int(self.node_off[n + 1])
                self._engine.write_occupancy(lo, np.array([occupancy_byte(it, u) for u in self.gpu_uuid[lo:hi]], dtype=np.uint8))
                return True
        return False

    # -- internals --------------------------------------------------------------------------------
    def _place(self, profile_names, lo, hi):
        req = np.zeros(len(profile_names), dtype=E.REQUEST_DTYPE)
        req["handle"] = np.arange(len(profile_names), dtype=np.uint32)
        req["profile"] = [self.profile_names.get(n, E.PROFILE_UNKNOWN) for n in profile_names]
        req["op"] = E.OP_ALLOC
        # ONE locked call restricts, places and restores (isl_place_batch_range): two reconcile workers cannot interleave and
        # nothing leaks when the call fails
        return self._engine.place_batch_range(lo, hi, req)

    def _release(self, res):
        spans = np.zeros(1, dtype=E.SPAN_DTYPE)
        spans[0] = (res["gpu"], res["start"], res["size"], 0)
        self._engine.free_batch(spans)

    def _details(self, instaslice, profileName, policy, pod, res):
        size, gi, ci, cieng = self.extractGpuProfile(instaslice, profileName)
        return policy.SetAllocationDetails(profileName, int(res["start"]), size, pod["uid"], instaslice["metadata"]["name"],
                                           "creating", gi, ci, cieng, pod.get("namespace", "default"), pod["name"],
                                           self.gpu_uuid[int(res["gpu"])])

    def _commit_or_veto(self, pod, profileName, policy, res):
        uuid = self.gpu_uuid[int(res["gpu"])]
        instaslice = self.items[self.node_of_uuid[uuid]]
        alloc = self._details(instaslice, profileName, policy, pod, res)
        for item in instaslice["spec"].get("prepared", {}).values():          # :198-203
            if item["parent"] == alloc["gpuUUID"] and item["size"] == alloc["size"] and item["start"] == alloc["start"]:
                self._release(res)
                return ("veto", None)
        instaslice["spec"].setdefault("allocations", {})[pod["uid"]] = alloc   # :215-219
        return ("placed", alloc)
