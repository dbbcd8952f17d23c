"""The latency path for a handful of pods (k_few: <= 8 requests, <= 16 384 GPUs, request-major block-wide search).  Needs a B200."""
import os

import numpy as np
import pytest

import oracle
from instaslice_b200 import engine as E
from instaslice_b200 import tables, workloads as W

pytestmark = pytest.mark.gpu


def tiny_batches(rng, ref, n_batches, n_profiles, max_n=8):
    """Batches of 1..max_n requests: ALLOCs (unknown profiles included), FREEs of live slices, a bad span, a NOOP now and then."""
    out, live = [], []
    for b in range(n_batches):
        n = 1 + int(rng.next1() % max_n)
        req = W.alloc_requests((rng.next(n) % np.uint64(n_profiles + 1)).astype(np.uint8))
        req["profile"][req["profile"] == n_profiles] = E.PROFILE_UNKNOWN
        for i in range(n):
            x = rng.next1() % 10
            if x < 3 and live:
                g, s, z = live.pop(int(rng.next1() % len(live)))
                req[i] = (g, 0, E.OP_FREE, s, z)
            elif x == 3:
                req[i] = (0xFFFFFF, 0, E.OP_FREE, 1, 1)            # bad span: GPU out of range
            elif x == 4:
                req[i] = (0, 0, E.OP_NOOP, 0, 0)
        res = ref.place(req)
        for r in res[(req["op"] == E.OP_ALLOC) & (res["status"] == E.ST_PLACED)]:
            live.append((int(r["gpu"]), int(r["start"]), int(r["size"])))
        out.append((req, res))
    return out


@pytest.mark.parametrize("quirks", [3, 0])
@pytest.mark.parametrize("G", [1, 37, 4096, 16384, 16400])
def test_tiny_batches_vs_oracle(G, quirks):
    rows = E.make_profiles(tables.H100_80GB)
    rng = W.SplitMix64(100 + G + quirks)
    node_off = np.concatenate([[0], np.cumsum(np.full((G + 7) // 8, 8))]).astype(np.uint32)
    node_off[-1] = G
    occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    if G > 100:
        occ[: G - 50] |= 0x7F                                   # nearly full: the search has to walk far
    ref = oracle.Fast(node_off, rows, quirks)
    ref.load(occ)
    batches = tiny_batches(rng, ref, 150, len(rows))
    for no_few in ("", "1"):                                    # k_few and, for the same calls, the fused k_small path
        os.environ.pop("ISL_NO_FEW", None)
        if no_few:
            os.environ["ISL_NO_FEW"] = "1"
        try:
            eng = E.Engine(max_gpus=max(4096, G), max_batch=1 << 12, quirks=quirks)
            eng.load_profiles(rows)
            eng.load_inventory(node_off, occ)
            for i, (req, want) in enumerate(batches):
                got = eng.place_batch(req)
                assert np.array_equal(got, want), (no_few, i, req, got, want)
            assert np.array_equal(eng.read_occupancy(), ref.occupancy())
            st = eng.stats()
            assert st["placed"] == sum(int((w["status"] == E.ST_PLACED).sum()) for r, w in batches)
        finally:
            os.environ.pop("ISL_NO_FEW", None)


def test_tiny_batches_heterogeneous_tables_and_partition():
    names, rows2d = E.make_profile_tables([tables.A100_40GB, tables.H100_80GB, tables.A30_24GB])
    rng = W.SplitMix64(7)
    n_nodes = 300
    node_off = W.node_offsets(n_nodes, 8)
    G = int(node_off[-1])
    node_table = (rng.next(n_nodes) % np.uint64(3)).astype(np.uint8)
    occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    occ[: G - 40] |= 0x7F
    ref = oracle.Fast(node_off, rows2d, 3, node_table=node_table)
    ref.load(occ)
    batches = tiny_batches(rng, ref, 120, len(names))
    eng = E.Engine(max_gpus=4096, max_batch=1 << 12)
    eng.load_profile_tables(rows2d)
    eng.load_inventory(node_off, occ)
    eng.set_node_tables(node_table)
    for i, (req, want) in enumerate(batches):
        assert np.array_equal(eng.place_batch(req), want), i
    assert np.array_equal(eng.read_occupancy(), ref.occupancy())
