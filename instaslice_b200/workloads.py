"""Synthetic inputs for the BASELINE configs (definitions: SURVEY.md section 8d).

Everything is driven by splitmix64 (seed 42 unless noted) so that any other implementation can emit the
identical streams.  The churn workload (config 4) needs to know where earlier pods were placed in order to
free them; it takes a ``placer`` callable (the engine in the bench, an oracle in CPU tests) and records the
batches it generated so they can be replayed.

  C1  1 node x 1 GPU, A100-40GB table, one ``1g.5gb`` request (samples/test-pod.yaml:16)
  C2  32 x 8 = 256 GPUs, 80GB-class table, 10 000 x ``1g.10gb``
  C3  512 x 8 = 4 096 GPUs, 100 000 requests, mix 1g 40 % / 2g 25 % / 3g 20 % / 4g 10 % / 7g 5 %
  C4  8 192 x 8 = 65 536 GPUs, pre-filled to 50 %, 1 000 000 ops in batches of 65 536, alloc/free 50/50
  C5  ``3g.20gb`` replay (samples/vllm_dep.yaml:36) on A100-40GB tables — see bench/latency tooling
"""
from __future__ import annotations

import numpy as np

from . import tables
from .engine import OP_ALLOC, OP_FREE, REQUEST_DTYPE, RESULT_DTYPE, ST_PLACED, make_profiles

_M64 = (1 << 64) - 1
_GOLD = 0x9E3779B97F4A7C15


class SplitMix64:
    """Vectorised splitmix64: ``next(n)`` returns the next n outputs as uint64."""

    def __init__(self, seed: int = 42):
        self.state = seed & _M64

    def next(self, n: int) -> np.ndarray:
        with np.errstate(over="ignore"):
            idx = np.arange(1, n + 1, dtype=np.uint64)
            z = np.uint64(self.state) + idx * np.uint64(_GOLD)
            self.state = (self.state + n * _GOLD) & _M64
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            return z ^ (z >> np.uint64(31))

    def next1(self) -> int:
        return int(self.next(1)[0])


def node_offsets(n_nodes: int, gpus_per_node: int = 8) -> np.ndarray:
    return (np.arange(n_nodes + 1, dtype=np.uint64) * gpus_per_node).astype(np.uint32)


# profile mix of configs 3/4 on the 80GB-class table (defined by this repo; the reference has none)
MIX_80GB = [("1g.10gb", 40), ("2g.20gb", 25), ("3g.40gb", 20), ("4g.40gb", 10), ("7g.80gb", 5)]


def mix_profiles(rng: SplitMix64, n: int, table=tables.H100_80GB, mix=MIX_80GB) -> np.ndarray:
    """n i.i.d. profile row indices drawn from ``mix`` (percent weights summing to 100)."""
    r = (rng.next(n) >> np.uint64(11)) % np.uint64(100)
    edges = np.cumsum([w for _, w in mix])
    rows = np.array([tables.profile_index(table, name) for name, _ in mix], dtype=np.uint8)
    return rows[np.searchsorted(edges, r.astype(np.int64), side="right")]


def alloc_requests(profiles: np.ndarray) -> np.ndarray:
    req = np.zeros(len(profiles), dtype=REQUEST_DTYPE)
    req["handle"] = np.arange(len(profiles), dtype=np.uint32)
    req["profile"] = profiles
    req["op"] = OP_ALLOC
    return req


def config1():
    """(node_off, occ, profile rows, requests) for BASELINE config 1."""
    rows = make_profiles(tables.A100_40GB)
    req = alloc_requests(np.array([tables.profile_index(tables.A100_40GB, "1g.5gb")], dtype=np.uint8))
    return node_offsets(1, 1), np.zeros(1, dtype=np.uint8), rows, req


def config2():
    rows = make_profiles(tables.H100_80GB)
    p = tables.profile_index(tables.H100_80GB, "1g.10gb")
    return node_offsets(32, 8), np.zeros(256, dtype=np.uint8), rows, alloc_requests(np.full(10_000, p, dtype=np.uint8))


def config3(seed: int = 42, n: int = 100_000, n_nodes: int = 512):
    rows = make_profiles(tables.H100_80GB)
    rng = SplitMix64(seed)
    return node_offsets(n_nodes, 8), np.zeros(n_nodes * 8, dtype=np.uint8), rows, alloc_requests(mix_profiles(rng, n))


class Churn:
    """BASELINE config 4: pre-fill to ``fill`` of the 7 usable slices per GPU, then ``n_ops`` operations in
    batches of ``batch``; each op is a FREE of a uniformly random live allocation with probability 1/2 (an
    ALLOC when nothing is live), else an ALLOC drawn from the mix.  Within a batch the engine applies the
    FREEs first, so a FREE may only name an allocation that was live when the batch started.

    ``placer(requests) -> results`` resolves one batch (engine or oracle); the generated batches are kept in
    ``self.batches`` (pre-fill batches first, ``self.n_prefill_batches`` of them) for replay.
    """

    def __init__(self, n_nodes: int = 8192, gpus_per_node: int = 8, n_ops: int = 1_000_000, batch: int = 65_536,
                 fill: float = 0.5, seed: int = 42, table=tables.H100_80GB, mix=MIX_80GB, min_age: int = 1):
        """``min_age`` = k: a FREE of churn batch b names an allocation placed by batch b - k or earlier (k = 1: any allocation live
        when the batch starts — the original definition).  With k > 1 a caller may legitimately have k batches in flight: it can
        compose batch b as soon as it has seen the results of batch b - k (the causal feed of ``isl_stream_submit``)."""
        assert min_age >= 1
        self.min_age = min_age
        self._young = []          # [(batch index, gpu[], start[], size[])] placements not yet old enough to be freed
        self.node_off = node_offsets(n_nodes, gpus_per_node)
        self.G = n_nodes * gpus_per_node
        self.rows = make_profiles(table)
        self.table, self.mix = table, mix
        self.n_ops, self.batch, self.fill = n_ops, batch, fill
        self.rng = SplitMix64(seed)
        self.batches: list[np.ndarray] = []
        self.n_prefill_batches = 0
        # live allocations as parallel arrays with swap-remove
        cap = self.G * 7 + batch * (min_age + 1)
        self._gpu = np.zeros(cap, dtype=np.uint32)
        self._start = np.zeros(cap, dtype=np.uint8)
        self._size = np.zeros(cap, dtype=np.uint8)
        self._live = 0
        self._busy_slices = 0

    def _absorb(self, req: np.ndarray, res: np.ndarray, batch_index: int | None = None):
        placed = (req["op"] == OP_ALLOC) & (res["status"] == ST_PLACED)
        self._busy_slices += int(res["size"][placed].astype(np.int64).sum())
        if batch_index is None:       # pre-fill: old enough from the start
            self._pool_add(res["gpu"][placed], res["start"][placed], res["size"][placed])
        else:
            self._young.append((batch_index, res["gpu"][placed].copy(), res["start"][placed].copy(), res["size"][placed].copy()))

    def _pool_add(self, gpu, start, size):
        k = len(gpu)
        self._gpu[self._live:self._live + k] = gpu
        self._start[self._live:self._live + k] = start
        self._size[self._live:self._live + k] = size
        self._live += k

    def _age(self, batch_index: int):
        """Placements of batches <= batch_index - min_age become eligible for FREEs (in placement order)."""
        while self._young and self._young[0][0] <= batch_index - self.min_age:
            _, gpu, start, size = self._young.pop(0)
            self._pool_add(gpu, start, size)

    def generate(self, placer, after_prefill=None):
        """Run pre-fill and churn through ``placer``; returns the list of recorded batches.
        ``after_prefill()`` is called once between the two phases (e.g. to snapshot the occupancy)."""
        target = self.fill * 7 * self.G
        prefill_batch = min(self.batch, max(64, self.G // 8))     # small steps so the target is not overshot by much
        while self._busy_slices < target:
            req = alloc_requests(mix_profiles(self.rng, prefill_batch, self.table, self.mix))
            res = placer(req)
            self._absorb(req, res)
            self.batches.append(req)
            self.n_prefill_batches += 1
            if self.n_prefill_batches > 256:
                raise RuntimeError("pre-fill did not reach the target occupancy")
        if after_prefill is not None:
            after_prefill()
        done = 0
        bi = 0
        while done < self.n_ops:
            self._age(bi)
            n = min(self.batch, self.n_ops - done)
            coin = (self.rng.next(n) >> np.uint64(63)).astype(bool)        # True -> FREE
            pick = self.rng.next(n)
            prof = mix_profiles(self.rng, n, self.table, self.mix)
            req = np.zeros(n, dtype=REQUEST_DTYPE)
            req["handle"] = np.arange(n, dtype=np.uint32)
            req["profile"] = prof
            req["op"] = OP_ALLOC
            live = self._live
            gpu, start, size = self._gpu, self._start, self._size
            for i in np.flatnonzero(coin):
                if live == 0:
                    continue                                                # nothing live -> stays an ALLOC
                j = int(pick[i] % np.uint64(live))
                req[i] = (gpu[j], 0, OP_FREE, start[j], size[j])
                self._busy_slices -= int(size[j])
                live -= 1
                gpu[j], start[j], size[j] = gpu[live], start[live], size[live]
            self._live = live
            res = placer(req)
            self._absorb(req, res, bi)
            self.batches.append(req)
            done += n
            bi += 1
        return self.batches


def as_results(n: int) -> np.ndarray:
    return np.zeros(n, dtype=RESULT_DTYPE)
