"""Speculative rounds inside the segment pipeline (isl_set_speculation): every inventory stage simulates its segment at
once from predicted queue heads and commits only once its entry is certified.  Whatever the predictions are worth, the results must be
byte-identical to the request-major CPU oracle (``oracle.Fast``) — single batches, streams with a causal window, open streams,
heterogeneous tables, the repaired quirk set, inventories that run full or stay empty.  Needs a B200."""
import numpy as np
import pytest

import oracle
from instaslice_b200 import engine as E
from instaslice_b200 import tables, workloads as W

pytestmark = pytest.mark.gpu


def _mixed(rng, n, rows):
    return W.alloc_requests(W.mix_profiles(rng, n))


def _check_stats(eng, want_chunks):
    st = eng.stats()
    assert st["spec_chunks"] == want_chunks, st
    assert st["spec_rounds"] >= want_chunks
    return st


@pytest.mark.parametrize("G,n,fill", [(4096, 100_000, 0x00), (65536, 65536, 0x7F), (65536, 40_000, 0x15), (1024, 9000, 0x33), (512, 5000, 0x00)])
def test_single_batch_speculative_equals_oracle(G, n, fill):
    rows = E.make_profiles(tables.H100_80GB)
    rng = W.SplitMix64(11 + G + n)
    node_off = W.node_offsets(G // 8, 8)
    occ = ((rng.next(G) & rng.next(G)) & np.uint64(fill)).astype(np.uint8)
    req = _mixed(rng, n, rows)
    ref = oracle.Fast(node_off, rows)
    ref.load(occ)
    want = ref.place(req)
    for mode, flags in ((E.SPEC_ON, E.FLAG_FORCE_PIPELINE), (E.SPEC_OFF, E.FLAG_FORCE_PIPELINE)):
        eng = E.Engine(max_gpus=G, max_batch=1 << 17, flags=flags)
        eng.set_speculation(mode)
        eng.load_profiles(rows)
        eng.load_inventory(node_off, occ)
        got = eng.place_batch(req)
        assert np.array_equal(got, want), (mode, int(np.argmax(got != want)))
        assert np.array_equal(eng.read_occupancy(), ref.occupancy())
        st = eng.stats()
        assert (st["spec_chunks"] >= 1) == (mode == E.SPEC_ON), st
        eng.close()


@pytest.mark.parametrize("window", [1, 2, 3])
def test_churn_stream_with_causal_window(window):
    """Config-4-shaped churn (FREEs of live allocations + mixed ALLOCs) as a device-side stream with a causal window: AUTO turns the
    speculative rounds on for windows 1..3."""
    rows = E.make_profiles(tables.H100_80GB)
    G, n, n_batches = 16384, 16384, 8
    rng = W.SplitMix64(300 + window)
    node_off = W.node_offsets(G // 8, 8)
    occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    ref = oracle.Fast(node_off, rows)
    ref.load(occ)
    live, batches, wants = [], [], []
    for b in range(n_batches):
        req = _mixed(rng, n, rows)
        k = min(len(live), n // 2)
        for _ in range(k):
            g, s, z = live.pop(int(rng.next1() % len(live)))
            req[int(rng.next1() % n)] = (g, 0, E.OP_FREE, s, z)
        res = ref.place(req)
        placed = res[(req["op"] == E.OP_ALLOC) & (res["status"] == E.ST_PLACED)]
        live.extend((int(r["gpu"]), int(r["start"]), int(r["size"])) for r in placed)
        batches.append(req); wants.append(res)
    eng = E.Engine(max_gpus=G, max_batch=n_batches * n)
    eng.load_profiles(rows)
    eng.load_inventory(node_off, occ)
    eng.set_causal_window(window)
    got = eng.place_stream(batches)
    for b in range(n_batches):
        assert np.array_equal(got[b], wants[b]), b
    assert np.array_equal(eng.read_occupancy(), ref.occupancy())
    _check_stats(eng, n_batches)
    eng.close()


def test_open_stream_speculative():
    rows = E.make_profiles(tables.H100_80GB)
    G, n, n_batches = 65536, 30000, 6
    rng = W.SplitMix64(77)
    node_off = W.node_offsets(G // 8, 8)
    occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    ref = oracle.Fast(node_off, rows)
    ref.load(occ)
    eng = E.Engine(max_gpus=G, max_batch=n_batches * 65536)
    eng.set_speculation(E.SPEC_ON)
    eng.load_profiles(rows)
    eng.load_inventory(node_off, occ)
    h_in = E.PinnedArray(n_batches * n, E.REQUEST_DTYPE)
    h_out = E.PinnedArray(n_batches * n, E.RESULT_DTYPE)
    eng.stream_open(n_batches)
    live = []
    for b in range(n_batches):          # strictly causal: batch b is composed from the results of batch b - 1
        req = _mixed(rng, n, rows)
        for _ in range(min(len(live), n // 2)):
            g, s, z = live.pop(int(rng.next1() % len(live)))
            req[int(rng.next1() % n)] = (g, 0, E.OP_FREE, s, z)
        h_in.array[b * n:(b + 1) * n] = req
        t = eng.stream_submit_ptr(n, h_in.ptr + 8 * b * n, h_out.ptr + 8 * b * n)
        eng.stream_wait(t)
        got = h_out.array[b * n:(b + 1) * n]
        want = ref.place(req)
        assert np.array_equal(got, want), b
        placed = got[(req["op"] == E.OP_ALLOC) & (got["status"] == E.ST_PLACED)]
        live.extend((int(r["gpu"]), int(r["start"]), int(r["size"])) for r in placed)
    eng.stream_close()
    assert np.array_equal(eng.read_occupancy(), ref.occupancy())
    _check_stats(eng, n_batches)
    h_in.free(); h_out.free()
    eng.close()


@pytest.mark.parametrize("quirks", [3, 0])
def test_speculative_with_two_tables_and_quirk_sets(quirks):
    """A30 + H100 nodes mixed, both quirk sets (FIXED: 3g at two starts, 7g places on empty GPUs): the group masses of the prediction are
    only heuristics there — the certification must still make every result exact."""
    names, rows_t = E.make_profile_tables([tables.H100_80GB, tables.A30_24GB])
    n_nodes, gpn = 1024, 8
    G = n_nodes * gpn
    rng = W.SplitMix64(4242)
    node_off = W.node_offsets(n_nodes, gpn)
    node_table = (rng.next(n_nodes) % np.uint64(3) == 0).astype(np.uint8)        # a third of the nodes are A30
    occ = ((rng.next(G) & rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    req = W.alloc_requests((rng.next(50_000) % np.uint64(len(names))).astype(np.uint8))
    ref = oracle.Fast(node_off, rows_t, quirks=quirks, node_table=node_table)
    ref.load(occ)
    want = ref.place(req)
    eng = E.Engine(max_gpus=G, max_batch=1 << 17, quirks=quirks, flags=E.FLAG_FORCE_PIPELINE)
    eng.set_speculation(E.SPEC_ON)
    eng.load_profile_tables(rows_t)
    eng.load_inventory(node_off, occ)
    eng.set_node_tables(node_table)
    got = eng.place_batch(req)
    assert np.array_equal(got, want), int(np.argmax(got != want))
    assert np.array_equal(eng.read_occupancy(), ref.occupancy())
    _check_stats(eng, 1)
    eng.close()


def test_many_speculative_calls_reuse_the_record_memory():
    """Records are validated by (call epoch, round) tags only: hundreds of calls over the same record memory, alternating shapes."""
    rows = E.make_profiles(tables.H100_80GB)
    G = 8192
    rng = W.SplitMix64(9)
    node_off = W.node_offsets(G // 8, 8)
    eng = E.Engine(max_gpus=G, max_batch=1 << 16, flags=E.FLAG_FORCE_PIPELINE)
    eng.set_speculation(E.SPEC_ON)
    eng.load_profiles(rows)
    ref = oracle.Fast(node_off, rows)
    for it in range(150):
        occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
        n = 2000 + int(rng.next1() % 9000)
        req = _mixed(rng, n, rows)
        ref.load(occ)
        eng.load_inventory(node_off, occ)
        assert np.array_equal(eng.place_batch(req), ref.place(req)), it
    eng.close()


@pytest.mark.parametrize("n_ranks,window", [(2, 1), (3, 2), (4, 1)])
def test_speculative_rounds_across_ranks_on_one_gpu(n_ranks, window):
    """A partitioned inventory: several engines in one process, each owning a GPU range, every engine's record memory wired into every
    other one (isl_connect_spec_local; CUDA IPC across processes).  The stages of all engines form ONE sequence and exchange their
    per-round records through each other's memory; PLACED records land in the owner's result array.  The owner's array alone == the
    global sequential first-fit."""
    import torch
    from instaslice_b200 import dist as D
    rows = E.make_profiles(tables.H100_80GB)
    rng = W.SplitMix64(999 + n_ranks)
    G = 4096
    node_off = W.node_offsets(G // 8, 8)
    occ0 = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    ref = oracle.Fast(node_off, rows)
    ref.load(occ0)
    batches, want, live = [], [], []
    for b in range(6):
        n = 3000 + 700 * b
        req = _mixed(rng, n, rows)
        for i in range(min(len(live), n // 3)):
            g, s, z = live.pop(int(rng.next1() % len(live)))
            req[int(rng.next1() % n)] = (g, 0, E.OP_FREE, s, z)
        res = ref.place(req)
        for r in res[(req["op"] == E.OP_ALLOC) & (res["status"] == E.ST_PLACED)]:
            live.append((int(r["gpu"]), int(r["start"]), int(r["size"])))
        batches.append(req); want.append(res)
    sizes = np.array([len(b) for b in batches], dtype=np.uint32)
    n_ops = int(sizes.sum())
    d_in = torch.from_numpy(np.concatenate(batches).view(np.int64).copy()).cuda()
    bounds = D.all_bounds(G, n_ranks, align=64)
    cuts = [lo for lo, _ in bounds] + [G]
    engines = []
    for r, (lo, hi) in enumerate(bounds):
        eng = E.Engine(max_gpus=G, max_batch=1 << 16)
        eng.load_profiles(rows)
        eng.load_inventory(node_off, occ0)
        eng.ipc_inbox_handle(); eng.ipc_spec_handle()          # allocate the shared buffers
        engines.append(eng)
    for r, eng in enumerate(engines):
        eng.connect_local(engines[r + 1] if r + 1 < n_ranks else None, has_prev=r > 0)
        eng.connect_owner_local(engines[0] if r > 0 else None)
        eng.set_ring_world(n_ranks)
        eng.connect_spec_local(n_ranks, r, engines, cuts)
        eng.set_causal_window(window)
        eng.set_speculation(E.SPEC_ON)
    torch.cuda.synchronize()
    for stream_id in (1, 2, 3):
        for eng, (lo, hi) in zip(engines, bounds):
            eng.load_inventory(node_off, occ0)
            eng.set_partition(lo, hi)
        for eng in engines:
            eng.place_stream_partitioned(sizes, d_in.data_ptr(), eng.device_results(), stream_id)
        for eng in engines:
            eng.synchronize()

        class _View:            # torch view of the owner's engine-owned result array (no copy)
            __cuda_array_interface__ = {"shape": (n_ops,), "typestr": "<i8", "data": (engines[0].device_results(), False), "version": 3}
        got = torch.as_tensor(_View(), device="cuda").cpu().numpy().view(E.RESULT_DTYPE)
        assert np.array_equal(got, np.concatenate(want)), (n_ranks, window, stream_id)
        occ = np.concatenate([eng.read_occupancy()[lo:hi] for eng, (lo, hi) in zip(engines, bounds)])
        assert np.array_equal(occ, ref.occupancy())
        st = engines[-1].stats()
        assert st["spec_chunks"] >= len(batches), st
    for eng in engines:
        eng.close()


@pytest.mark.parametrize("tname", ["a100-40gb", "h100-80gb", "a30-24gb", "b200-180gb"])
@pytest.mark.parametrize("quirks", [3, 0])
def test_randomised_inventories_and_frees_speculative(tname, quirks):
    """Ragged nodes, occupancy with the unusable slot set, unknown profiles, frees of live allocations, batches of 1..6000 requests on a few
    hundred to a few thousand GPUs — with the rounds forced on: entries that are idle in one round and not in the next, windows that are
    re-used or staged anew, simulations that are cut off.  Everything byte-identical to the oracle."""
    table = tables.TABLES[tname]
    rows = E.make_profiles(table)
    rng = W.SplitMix64(4711 + quirks + len(tname))
    import os
    for trial in range(6 * int(os.environ.get("ISL_STRESS", "1"))):
        n_nodes = 8 + int(rng.next1() % 900)
        node_off = np.concatenate([[0], np.cumsum(1 + (rng.next(n_nodes) % np.uint64(9)).astype(np.int64))]).astype(np.uint32)
        G = int(node_off[-1])
        occ = ((rng.next(G) & rng.next(G)) & np.uint64(0xFF)).astype(np.uint8)
        eng = E.Engine(max_gpus=max(4096, G), max_batch=1 << 16, quirks=quirks, flags=E.FLAG_FORCE_PIPELINE | E.FLAG_NO_SMALL)
        eng.set_speculation(E.SPEC_ON)
        eng.load_profiles(rows)
        eng.load_inventory(node_off, occ)
        ref = oracle.Fast(node_off, rows, quirks)
        ref.load(occ)
        live = []
        for batch in range(5):
            n = 1 + int(rng.next1() % 6000)
            req = W.alloc_requests((rng.next(n) % np.uint64(len(table) + 1)).astype(np.uint8))
            req["profile"][req["profile"] == len(table)] = E.PROFILE_UNKNOWN
            for i in range(min(len(live), n // 3)):
                g, s, z = live.pop(int(rng.next1() % len(live)))
                req[int(rng.next1() % n)] = (g, 0, E.OP_FREE, s, z)
            got, want = eng.place_batch(req), ref.place(req)
            bad = np.flatnonzero(got != want)
            assert len(bad) == 0, (tname, quirks, trial, batch, bad[:4], got[bad[:4]], want[bad[:4]])
            assert np.array_equal(eng.read_occupancy(), ref.occupancy())
            for r in got[(req["op"] == E.OP_ALLOC) & (got["status"] == E.ST_PLACED)]:
                live.append((int(r["gpu"]), int(r["start"]), int(r["size"])))
        eng.close()
