"""Heterogeneous clusters: every node publishes its own Migplacement (instaslice_daemonset.go:588-664) and the reference
looks a profile up in the table of the node it is scanning (instaslice_controller.go:332-340).  Needs a B200."""
import copy

import numpy as np
import pytest

import oracle
from oracle import ref_py
from instaslice_b200 import controller as ctl
from instaslice_b200 import engine as E
from instaslice_b200 import tables, workloads as W

pytestmark = pytest.mark.gpu


def mixed(rng, n_nodes, gpus_per_node=8):
    names, rows2d = E.make_profile_tables([tables.A100_40GB, tables.H100_80GB])
    node_table = (rng.next(n_nodes) % np.uint64(2)).astype(np.uint8)
    node_off = W.node_offsets(n_nodes, gpus_per_node)
    return names, rows2d, node_table, node_off


@pytest.mark.parametrize("quirks", [3, 0])
@pytest.mark.parametrize("n_nodes", [3, 600, 2300])
def test_mixed_tables_vs_oracle(n_nodes, quirks):
    rng = W.SplitMix64(n_nodes + quirks)
    names, rows2d, node_table, node_off = mixed(rng, n_nodes)
    G = int(node_off[-1])
    occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    ref = oracle.Fast(node_off, rows2d, quirks, node_table=node_table)
    ref.load(occ)
    batches, live = [], []
    for b in range(3):
        n = 1 + int(rng.next1() % (4 * G))
        req = W.alloc_requests((rng.next(n) % np.uint64(len(names) + 1)).astype(np.uint8))
        req["profile"][req["profile"] == len(names)] = E.PROFILE_UNKNOWN
        for i in range(min(len(live), n // 3)):
            g, s, z = live.pop(int(rng.next1() % len(live)))
            req[int(rng.next1() % n)] = (g, 0, E.OP_FREE, s, z)
        batches.append(req)
        res = ref.place(req)
        for r in res[(req["op"] == E.OP_ALLOC) & (res["status"] == E.ST_PLACED)]:
            live.append((int(r["gpu"]), int(r["start"]), int(r["size"])))
    ref.load(occ)
    want = [ref.place(b) for b in batches]
    final = ref.occupancy()
    for flags in (E.FLAG_NO_PIPELINE | E.FLAG_NO_SMALL, E.FLAG_FORCE_PIPELINE, 0):
        eng = E.Engine(max_gpus=max(4096, G), max_batch=1 << 20, quirks=quirks, flags=flags)
        eng.load_profile_tables(rows2d)
        eng.load_inventory(node_off, occ)
        eng.set_node_tables(node_table)
        got = eng.place_stream(batches) if flags == 0 else [eng.place_batch(b) for b in batches]
        for i, (g_, w) in enumerate(zip(got, want)):
            bad = np.flatnonzero(g_ != w)
            assert len(bad) == 0, (flags, i, bad[:5], g_[bad[:5]], w[bad[:5]], batches[i][bad[:5]])
        assert np.array_equal(eng.read_occupancy(), final), flags
    # the device table of both tables: "1g.10gb" is a 2-slice profile on A100-40GB and a 1-slice profile on H100-80GB
    p = names.index("1g.10gb")
    occ_all = np.arange(256, dtype=np.uint8)
    for t in range(2):
        want_t = np.array([oracle.start_for(rows2d[t, p], quirks, o) for o in range(256)], dtype=np.uint8)
        assert np.array_equal(eng.eval_starts(p | (t << 8), occ_all), want_t)


def test_mixed_cluster_through_the_mirror_vs_python_restatement():
    rng = W.SplitMix64(77)
    tabs = [tables.A100_40GB, tables.H100_80GB]
    items = []
    g = 0
    for n in range(7):
        t = int(rng.next1() % 2)
        spec = {"MigGPUUUID": {}, "allocations": {}, "prepared": {}, "migplacement": tables.migplacement(tabs[t])}
        for _ in range(1 + int(rng.next1() % 3)):
            uuid = "GPU-%012d" % g
            g += 1
            spec["MigGPUUUID"][uuid] = "x"
            if rng.next1() % 2:
                spec["prepared"]["MIG-%d" % g] = {"profile": "", "start": int(rng.next1() % 3), "size": 2, "parent": uuid, "podUUID": "", "giinfo": 0, "ciinfo": 0}
        items.append({"metadata": {"name": "node-%02d" % n}, "spec": spec})
    names = sorted({row[0] for tab in tabs for row in tab})
    pods = [{"uid": "u%d" % k, "name": "p%d" % k, "profile": names[int(rng.next1() % len(names))]} for k in range(60)]
    want_items = copy.deepcopy(items)
    want = [ref_py.reconcile_gated_pod(want_items, p, p["profile"]) for p in pods]
    r = ctl.InstasliceReconciler(items)
    got = r.place_pending_pods(pods)
    for k, ((gv, ga), (wv, wa)) in enumerate(zip(got, want)):
        assert gv == wv, (k, pods[k])
        if ga:
            assert (ga["gpuUUID"], ga["nodename"], ga["start"], ga["size"], ga["giprofileid"]) == \
                   (wa[0]["gpuUUID"], wa[0]["nodename"], wa[0]["start"], wa[0]["size"], wa[0]["giprofileid"]), k
    # per-node lookup through the reference-named helper: same name, different answer on different node types
    for it in items:
        uuid = sorted(it["spec"]["MigGPUUUID"])[0]
        for name in ("1g.10gb", "1g.5gb", "3g.40gb"):
            assert r.getStartIndexFromPreparedState(it, uuid, name) == ref_py.get_start_index_from_prepared_state(it, uuid, name)
