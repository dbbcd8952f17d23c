"""Single-profile chunks (parallel capacity-scan commit) and the host-carried token API.  Needs a B200."""
import numpy as np
import pytest

import oracle
from instaslice_b200 import engine as E
from instaslice_b200 import tables, workloads as W

pytestmark = pytest.mark.gpu
FLAGS = E.FLAG_NO_PIPELINE | E.FLAG_NO_SMALL          # the multi-kernel path, where scan mode lives


@pytest.mark.parametrize("quirks", [3, 0])
@pytest.mark.parametrize("profile", ["1g.10gb", "2g.20gb", "3g.40gb", "7g.80gb"])
def test_single_profile_chunks_use_the_scan_and_match_the_oracle(profile, quirks):
    rows = E.make_profiles(tables.H100_80GB)
    p = tables.profile_index(tables.H100_80GB, profile)
    rng = W.SplitMix64(p * 7 + quirks)
    G = 9000
    node_off = np.concatenate([[0], np.cumsum(np.full((G + 7) // 8, 8))]).astype(np.uint32)
    node_off[-1] = G
    occ = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    ref = oracle.Fast(node_off, rows, quirks)
    ref.load(occ)
    eng = E.Engine(max_gpus=1 << 14, max_batch=1 << 18, quirks=quirks, flags=FLAGS)
    eng.load_profiles(rows)
    eng.load_inventory(node_off, occ)
    live = []
    for n in (5000, 70000, 3):           # 70000 = two chunks of one call, each single-profile
        req = W.alloc_requests(np.full(n, p, dtype=np.uint8))
        req["profile"][::97] = E.PROFILE_UNKNOWN                  # unknown profiles do not disturb scan mode
        for i in range(min(len(live), n // 4)):
            g, s, z = live.pop(int(rng.next1() % len(live)))
            req[int(rng.next1() % n)] = (g, 0, E.OP_FREE, s, z)
        got, want = eng.place_batch(req), ref.place(req)
        bad = np.flatnonzero(got != want)
        assert len(bad) == 0, (n, bad[:5], got[bad[:5]], want[bad[:5]])
        assert np.array_equal(eng.read_occupancy(), ref.occupancy())
        for r in got[(req["op"] == E.OP_ALLOC) & (got["status"] == E.ST_PLACED)]:
            live.append((int(r["gpu"]), int(r["start"]), int(r["size"])))
    st = eng.stats()
    assert st["chain_steps"] == 0 and st["scan_placed"] == st["placed"]       # everything went through the scan, nothing through the chain


def test_config2_is_committed_by_the_scan():
    node_off, occ, rows, req = W.config2()
    eng = E.Engine(max_gpus=4096, max_batch=1 << 16)
    eng.load_profiles(rows)
    eng.load_inventory(node_off, occ)
    res = eng.place_batch(req)
    k = np.arange(len(req))
    placed = k < 1792
    assert np.array_equal(res["status"] == E.ST_PLACED, placed)
    assert np.array_equal(res["gpu"][placed], (k[placed] // 7).astype(np.uint32)) and np.array_equal(res["start"][placed], (k[placed] % 7).astype(np.uint8))
    st = eng.stats()
    assert st["scan_placed"] == 1792 and st["chain_steps"] == 0


@pytest.mark.parametrize("single_profile", [False, True])
def test_host_carried_token_between_two_engines(single_profile):
    """isl_place_batch_partitioned: the caller carries the queue-head token from the engine that owns the lower GPU range
    to the next one (device buffers); merged results == global first-fit.  Covers chain mode and scan mode."""
    import torch
    rows = E.make_profiles(tables.H100_80GB)
    rng = W.SplitMix64(5 + single_profile)
    G = 6000
    node_off = np.concatenate([[0], np.cumsum(np.full((G + 7) // 8, 8))]).astype(np.uint32)
    node_off[-1] = G
    occ = ((rng.next(G) | rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
    n = 70000
    prof = np.full(n, 0, dtype=np.uint8) if single_profile else W.mix_profiles(rng, n)
    req = W.alloc_requests(prof)
    ref = oracle.Fast(node_off, rows)
    ref.load(occ)
    want = ref.place(req)
    d_in = torch.from_numpy(req.view(np.int64).copy()).cuda()
    cut = 2500
    outs, heads = [], torch.zeros(2 * 16, dtype=torch.int32, device="cuda")
    engines = []
    for r, (lo, hi) in enumerate(((0, cut), (cut, G))):
        eng = E.Engine(max_gpus=1 << 13, max_batch=1 << 17, flags=FLAGS)
        eng.load_profiles(rows)
        eng.load_inventory(node_off, occ)
        eng.set_partition(lo, hi)
        out = torch.empty_like(d_in)
        nxt = torch.zeros_like(heads)
        eng.place_batch_partitioned(n, d_in.data_ptr(), out.data_ptr(), heads.data_ptr() if r else None, nxt.data_ptr())
        eng.synchronize()
        heads = nxt
        outs.append(out.cpu().numpy())
        engines.append((eng, lo, hi))
    merged = np.minimum(outs[0], outs[1]).view(E.RESULT_DTYPE)
    assert np.array_equal(merged, want)
    occ_got = np.concatenate([eng.read_occupancy()[lo:hi] for eng, lo, hi in engines])
    assert np.array_equal(occ_got, ref.occupancy())
