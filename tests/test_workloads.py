"""Generators of the BASELINE configs (CPU only; the oracle plays the placer)."""
import numpy as np

import oracle
from instaslice_b200 import engine as E
from instaslice_b200 import tables, workloads as W


def test_splitmix64_reference_values():
    # first outputs of splitmix64 seeded with 0 (public reference values of the algorithm)
    r = W.SplitMix64(0).next(3)
    assert [int(x) for x in r] == [0xE220A8397B1DCDAF, 0x6E789E6AA1B965F4, 0x06C45D188009454F]
    a, b = W.SplitMix64(42), W.SplitMix64(42)
    x = a.next(10)
    y = np.concatenate([b.next(4), b.next(6)])
    assert np.array_equal(x, y)          # chunking does not change the stream


def test_mix_proportions():
    prof = W.mix_profiles(W.SplitMix64(1), 200_000)
    names = [tables.H100_80GB[p][0] for p in range(len(tables.H100_80GB))]
    frac = {names[p]: float((prof == p).mean()) for p in range(len(names))}
    for name, pct in W.MIX_80GB:
        assert abs(frac[name] - pct / 100) < 0.01
    assert frac["1g.20gb"] == 0.0


def test_churn_small_is_consistent():
    ch = W.Churn(n_nodes=256, gpus_per_node=8, n_ops=4000, batch=1024, seed=5)
    ref = oracle.Fast(ch.node_off, ch.rows)
    ref.load(np.zeros(ch.G, dtype=np.uint8))
    batches = ch.generate(ref.place)
    assert ch.n_prefill_batches >= 1 and len(batches) == ch.n_prefill_batches + 4
    # replay on a fresh oracle: every FREE names a span that is fully busy at the time it is applied
    ref2 = oracle.Fast(ch.node_off, ch.rows)
    ref2.load(np.zeros(ch.G, dtype=np.uint8))
    n_free = 0
    for req in batches:
        occ = ref2.occupancy()
        fr = req[req["op"] == E.OP_FREE]
        n_free += len(fr)
        for g, s, z in zip(fr["handle"], fr["start"], fr["size"]):
            span = ((1 << int(z)) - 1) << int(s)
            assert occ[g] & span == span
            occ[g] &= ~span & 0xFF                 # no span is freed twice within a batch
        ref2.place(req)
    assert n_free > 1000
    # occupancy hovers around the pre-fill target
    busy = int(np.unpackbits(ref2.occupancy()).sum())
    assert 0.35 * 7 * ch.G < busy < 0.65 * 7 * ch.G


def test_churn_min_age_is_causal():
    """min_age = k: every FREE of churn batch b names an allocation placed by churn batch b - k or earlier (or by the pre-fill), i.e.
    a caller that has seen the results of batch b - k can compose batch b (the causal feed of isl_stream_submit); k = 1 is the
    original stream."""
    for k in (1, 3):
        ch = W.Churn(n_nodes=256, gpus_per_node=8, n_ops=8000, batch=1024, seed=5, min_age=k)
        ref = oracle.Fast(ch.node_off, ch.rows)
        ref.load(np.zeros(ch.G, dtype=np.uint8))
        placed_by = {}                      # (gpu, start, size) -> churn batch that placed the live allocation (-1: pre-fill)
        state = {"churn": False, "b": 0, "frees": 0}

        def placer(req):
            b = state["b"] if state["churn"] else -(1 << 20)      # the pre-fill happened long ago
            fr = req[req["op"] == E.OP_FREE]
            for g, s, z in zip(fr["handle"].tolist(), fr["start"].tolist(), fr["size"].tolist()):
                assert placed_by.pop((g, s, z)) <= b - k, (k, b)
                state["frees"] += 1
            res = ref.place(req)
            for r in res[(req["op"] == E.OP_ALLOC) & (res["status"] == E.ST_PLACED)]:
                placed_by[(int(r["gpu"]), int(r["start"]), int(r["size"]))] = b
            if state["churn"]:
                state["b"] += 1
            return res

        ch.generate(placer, after_prefill=lambda: state.update(churn=True))
        assert state["frees"] > 1000 and state["b"] == 8
