"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / synccheck): both device paths, frees, stream, best-fit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from instaslice_b200 import engine as E, tables, workloads as W
rows = E.make_profiles(tables.H100_80GB)
rng = W.SplitMix64(5)
G = 2048
node_off = W.node_offsets(G // 8, 8)
occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
ref = oracle.Fast(node_off, rows); ref.load(occ)
batches, live = [], []
for b in range(3):
    n = 1500
    req = W.alloc_requests(W.mix_profiles(rng, n))
    for i in range(min(len(live), n // 3)):
        g, s, z = live.pop(int(rng.next1() % len(live)))
        req[int(rng.next1() % n)] = (g, 0, E.OP_FREE, s, z)
    res = ref.place(req)
    for r in res[(req["op"] == E.OP_ALLOC) & (res["status"] == E.ST_PLACED)]:
        live.append((int(r["gpu"]), int(r["start"]), int(r["size"])))
    batches.append((req, res))
for flags in (E.FLAG_NO_PIPELINE, E.FLAG_FORCE_PIPELINE):
    eng = E.Engine(max_gpus=4096, max_batch=1 << 16, flags=flags)
    eng.load_profiles(rows); eng.load_inventory(node_off, occ)
    for req, res in batches:
        assert np.array_equal(eng.place_batch(req), res)
eng = E.Engine(max_gpus=4096, max_batch=1 << 16)
eng.load_profiles(rows); eng.load_inventory(node_off, occ)
got = eng.place_stream([b[0] for b in batches])
assert all(np.array_equal(g, b[1]) for g, b in zip(got, batches))
# speculative rounds: a stream with one batch in flight, and forced on for the unconstrained stream
for window, mode in ((1, E.SPEC_AUTO), (0, E.SPEC_ON)):
    sp = E.Engine(max_gpus=4096, max_batch=1 << 16)
    sp.set_speculation(mode); sp.set_causal_window(window)
    sp.load_profiles(rows); sp.load_inventory(node_off, occ)
    got = sp.place_stream([b[0] for b in batches])
    assert all(np.array_equal(g, b[1]) for g, b in zip(got, batches)) and sp.stats()["spec_chunks"] == len(batches)
bf = E.Engine(max_gpus=4096, max_batch=1 << 16, policy=E.POLICY_BEST_FIT)
bf.load_profiles(rows); bf.load_inventory(node_off, occ)
rb = oracle.Fast(node_off, rows, 3, policy=1); rb.load(occ)
assert np.array_equal(bf.place_batch(batches[0][0]), rb.place(batches[0][0]))
# fused small-batch kernel, scan mode, heterogeneous tables
small = E.Engine(max_gpus=4096, max_batch=1 << 16)
small.load_profiles(rows); small.load_inventory(node_off, occ)
rs = oracle.Fast(node_off, rows); rs.load(occ)
for n in (1, 40, 700):
    rq = W.alloc_requests(W.mix_profiles(rng, n))
    assert np.array_equal(small.place_batch(rq), rs.place(rq))
scan = E.Engine(max_gpus=4096, max_batch=1 << 16, flags=E.FLAG_NO_PIPELINE | E.FLAG_NO_SMALL)
scan.load_profiles(rows); scan.load_inventory(node_off, occ)
rs.load(occ)
rq = W.alloc_requests(np.zeros(5000, dtype=np.uint8))
assert np.array_equal(scan.place_batch(rq), rs.place(rq)) and scan.stats()["scan_placed"] > 0
names, rows2d = E.make_profile_tables([tables.A100_40GB, tables.H100_80GB])
nt = (rng.next(G // 8) % np.uint64(2)).astype(np.uint8)
rh = oracle.Fast(node_off, rows2d, 3, node_table=nt); rh.load(occ)
bh = [W.alloc_requests((rng.next(n) % np.uint64(len(names))).astype(np.uint8)) for n in (3000, 900, 70000)]
wh = [rh.place(b) for b in bh]
for flags in (E.FLAG_NO_PIPELINE | E.FLAG_NO_SMALL, 0):
    he = E.Engine(max_gpus=4096, max_batch=1 << 17, flags=flags)
    he.load_profile_tables(rows2d); he.load_inventory(node_off, occ); he.set_node_tables(nt)
    gh = he.place_stream(bh) if flags == 0 else [he.place_batch(b) for b in bh]
    assert all(np.array_equal(a, b) for a, b in zip(gh, wh))
print("sanitize run ok")
