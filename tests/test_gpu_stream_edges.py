"""Edge cases of the stream / segment-pipeline path.  Needs a B200."""
import numpy as np
import pytest

import oracle
from instaslice_b200 import engine as E
from instaslice_b200 import tables, workloads as W

pytestmark = pytest.mark.gpu


def churn_batches(rng, ref, sizes, n_profiles):
    batches, live = [], []
    for n in sizes:
        req = W.alloc_requests((rng.next(n) % np.uint64(n_profiles)).astype(np.uint8)) if n else np.zeros(0, dtype=E.REQUEST_DTYPE)
        for i in range(min(len(live), n // 3)):
            g, s, z = live.pop(int(rng.next1() % len(live)))
            req[int(rng.next1() % n)] = (g, 0, E.OP_FREE, s, z)
        res = ref.place(req)
        for r in res[(req["op"] == E.OP_ALLOC) & (res["status"] == E.ST_PLACED)]:
            live.append((int(r["gpu"]), int(r["start"]), int(r["size"])))
        batches.append((req, res))
    return batches


def test_stream_with_empty_and_tiny_batches():
    """300 batches of 0..40 requests (empty ones included) in ONE call: one pipeline chunk per non-empty batch."""
    rows = E.make_profiles(tables.H100_80GB)
    rng = W.SplitMix64(11)
    G = 2048
    node_off = W.node_offsets(G // 8, 8)
    occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    ref = oracle.Fast(node_off, rows)
    ref.load(occ)
    sizes = [int(x) for x in (rng.next(300) % np.uint64(41))]
    sizes[0] = 0
    sizes[-1] = 0
    batches = churn_batches(rng, ref, sizes, len(rows))
    for flags in (0, E.FLAG_NO_PIPELINE | E.FLAG_NO_SMALL):
        eng = E.Engine(max_gpus=4096, max_batch=1 << 16, flags=flags)
        eng.load_profiles(rows)
        eng.load_inventory(node_off, occ)
        got = eng.place_stream([b[0] for b in batches])
        assert len(got) == len(batches)
        for i, (g_, (_, w)) in enumerate(zip(got, batches)):
            assert np.array_equal(g_, w), (flags, i)
        assert np.array_equal(eng.read_occupancy(), ref.occupancy())
    # a stream made only of empty batches is a no-op
    assert all(len(x) == 0 for x in eng.place_stream([np.zeros(0, dtype=E.REQUEST_DTYPE)] * 3))


def test_engine_reuse_across_inventories_and_tables():
    """One engine, reloaded with inventories of different size and different tables between stream calls."""
    eng = E.Engine(max_gpus=1 << 16, max_batch=1 << 18)
    rng = W.SplitMix64(21)
    for G, table in ((65536, tables.H100_80GB), (512, tables.A100_40GB), (20000, tables.H100_80GB), (8, tables.A100_40GB)):
        rows = E.make_profiles(table)
        node_off = np.concatenate([[0], np.cumsum(np.full((G + 7) // 8, 8))]).astype(np.uint32)
        node_off[-1] = G
        occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
        ref = oracle.Fast(node_off, rows)
        ref.load(occ)
        batches = churn_batches(rng, ref, [5000, 70000, 1, 3000], len(rows))
        eng.load_profiles(rows)
        eng.load_inventory(node_off, occ)
        got = eng.place_stream([b[0] for b in batches])
        for i, (g_, (_, w)) in enumerate(zip(got, batches)):
            assert np.array_equal(g_, w), (G, i)
        assert np.array_equal(eng.read_occupancy(), ref.occupancy()), G


def test_stream_capacity_errors():
    rows = E.make_profiles(tables.H100_80GB)
    eng = E.Engine(max_gpus=4096, max_batch=1000)
    eng.load_profiles(rows)
    eng.load_inventory(W.node_offsets(4, 8), np.zeros(32, dtype=np.uint8))
    with pytest.raises(E.EngineError) as ei:
        eng.place_stream([W.alloc_requests(np.zeros(600, dtype=np.uint8)), W.alloc_requests(np.zeros(600, dtype=np.uint8))])
    assert ei.value.code == E.ERANGE
    ok = eng.place_stream([W.alloc_requests(np.zeros(500, dtype=np.uint8)), W.alloc_requests(np.zeros(500, dtype=np.uint8))])
    assert sum(int((r["status"] == E.ST_PLACED).sum()) for r in ok) == 32 * 7


@pytest.mark.parametrize("sizes", [[70001, 0, 513, 65536, 3, 99999], [1500, 1500], [65537, 65535, 1]])
def test_stream_from_pinned_host_buffers(sizes):
    """isl_place_stream with pinned host buffers: the batches are fed while the pipeline runs and an extra CTA writes every finished
    chunk into the caller's array (odd offsets, empty batches, multi-chunk batches); same answers as pageable buffers and the oracle."""
    import torch
    rows = E.make_profiles(tables.H100_80GB)
    rng = W.SplitMix64(31 + len(sizes))
    G = 30000
    node_off = np.concatenate([[0], np.cumsum(np.full((G + 7) // 8, 8))]).astype(np.uint32)
    node_off[-1] = G
    occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    ref = oracle.Fast(node_off, rows)
    ref.load(occ)
    batches = churn_batches(rng, ref, sizes, len(rows))
    want = np.concatenate([b[1] for b in batches])
    req = np.concatenate([b[0] for b in batches])
    h_in = torch.from_numpy(req.view(np.uint64).copy()).pin_memory()
    h_out = torch.zeros_like(h_in).pin_memory()
    eng = E.Engine(max_gpus=1 << 15, max_batch=1 << 19)
    eng.load_profiles(rows)
    for rep in range(2):        # twice: flags and counters of the first call must not leak into the second
        eng.load_inventory(node_off, occ)
        h_out.zero_()
        eng.place_stream_ptr(np.array(sizes, dtype=np.uint32), h_in.data_ptr(), h_out.data_ptr(), device=False)
        got = h_out.numpy().view(E.RESULT_DTYPE)
        bad = np.flatnonzero(got != want)
        assert len(bad) == 0, (rep, bad[:5], got[bad[:5]], want[bad[:5]])
        assert np.array_equal(eng.read_occupancy(), ref.occupancy())
    eng.load_inventory(node_off, occ)
    pageable = eng.place_stream([b[0] for b in batches])
    assert np.array_equal(np.concatenate(pageable), want)


def test_what_if_snapshot_and_restore():
    """SURVEY 8f-4 what-if queries: snapshot, try a placement plan (frees + allocs), restore — the live state is untouched,
    and a snapshot does not survive a new inventory."""
    rows = E.make_profiles(tables.H100_80GB)
    rng = W.SplitMix64(41)
    G = 5000
    node_off = np.concatenate([[0], np.cumsum(np.full((G + 7) // 8, 8))]).astype(np.uint32)
    node_off[-1] = G
    occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    eng = E.Engine(max_gpus=1 << 13, max_batch=1 << 18)
    eng.load_profiles(rows)
    with pytest.raises(E.EngineError):
        eng.restore_occupancy()                       # nothing loaded, nothing snapshotted
    eng.load_inventory(node_off, occ)
    with pytest.raises(E.EngineError):
        eng.restore_occupancy()
    eng.snapshot_occupancy()
    plan = [W.alloc_requests(W.mix_profiles(rng, n)) for n in (3000, 70000)]
    first = eng.place_stream(plan)
    assert not np.array_equal(eng.read_occupancy(), occ)
    eng.restore_occupancy()
    assert np.array_equal(eng.read_occupancy(), occ)
    again = eng.place_stream(plan)                    # the same question gets the same answer
    assert all(np.array_equal(a, b) for a, b in zip(first, again))
    eng.load_inventory(node_off, occ)
    with pytest.raises(E.EngineError):
        eng.restore_occupancy()


def test_token_tag_wraps_after_32768_stream_calls():
    """The token words of the segment pipeline carry the low 15 bits of the call epoch as their validity tag.  33 000 stream calls on one
    engine cross the wrap-around (tag 0 is skipped, stale tags of 32 768 calls ago are cleared): every call must answer like the first."""
    rows = E.make_profiles(tables.H100_80GB)
    G = 1024                                   # two segments: tokens really travel
    node_off = np.concatenate([[0], np.cumsum(np.full(G // 8, 8))]).astype(np.uint32)
    rng = W.SplitMix64(51)
    occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    eng = E.Engine(max_gpus=G, max_batch=1 << 12, flags=E.FLAG_FORCE_PIPELINE)
    eng.load_profiles(rows)
    eng.load_inventory(node_off, occ)
    b1 = W.alloc_requests(np.array([0, 2, 3, 0, 1, 4], dtype=np.uint8))
    first = eng.place_stream([b1, b1[:1]])
    placed = np.concatenate(first)
    placed = placed[placed["status"] == E.ST_PLACED]
    undo = np.zeros(len(placed), dtype=E.REQUEST_DTYPE)
    undo["handle"], undo["op"], undo["start"], undo["size"] = placed["gpu"], E.OP_FREE, placed["start"], placed["size"]
    ref = oracle.Fast(node_off, rows)
    ref.load(occ)
    assert np.array_equal(first[0], ref.place(b1))
    eng.place_stream([undo, undo[:0]])            # back to the loaded occupancy
    assert np.array_equal(eng.read_occupancy(), occ)
    sizes = np.array([len(b1), 1, len(undo)], dtype=np.uint32)
    req = np.concatenate([b1, b1[:1], undo])
    out = np.empty(len(req), dtype=E.RESULT_DTYPE)
    want = None
    for call in range(33000):
        eng.place_stream_ptr(sizes, req.ctypes.data, out.ctypes.data, device=False)
        if want is None:
            want = out.copy()
            assert np.array_equal(want[:len(b1)], first[0]) and np.array_equal(want[len(b1):len(b1) + 1], first[1])
        elif not np.array_equal(out, want):
            raise AssertionError("call %d differs" % call)
    assert np.array_equal(eng.read_occupancy(), occ)
