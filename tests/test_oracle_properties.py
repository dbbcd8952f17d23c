"""Property tests of the oracle (hypothesis, CPU only): structural facts the reference's search implies, independent of any
particular vector — they guard the restatements against each other and against the single rule `start_for`."""
import numpy as np
from hypothesis import given, settings, strategies as st

import oracle
from instaslice_b200 import engine as E
from instaslice_b200.workloads import alloc_requests

row_st = st.builds(lambda size, starts, gi: ("p", size, starts, gi),
                   st.integers(1, 8), st.lists(st.integers(0, 7), min_size=1, max_size=8, unique=True), st.integers(0, 20))


@settings(max_examples=300, deadline=None)
@given(row=row_st, occ=st.integers(0, 255), quirks=st.sampled_from([0, 1, 2, 3]))
def test_start_is_legal_free_and_first(row, occ, quirks):
    r = E.make_profiles([row])[0]
    s = oracle.start_for(r, quirks, occ)
    size, starts = row[1], row[2]
    strict, pow2 = quirks & 1, quirks & 2

    def ok(v):
        if (occ >> v) & 1:
            return False
        if size == 1:
            return True
        if pow2 and size not in (2, 4, 8):
            return False
        if (v + size >= 8) if strict else (v + size > 8):
            return False
        return occ & (((1 << size) - 1) << v) == 0

    legal = [v for v in starts if ok(v)]
    assert s == (legal[0] if legal else 9)            # the first legal start in ROW order, or the sentinel 9


@settings(max_examples=60, deadline=None)
@given(data=st.data())
def test_first_fit_is_monotone_and_never_double_books(data):
    """Within one ALLOC phase: occupancy only grows, no two placements overlap, and a request is only refused
    when no GPU can take it at that moment (checked against the per-byte rule)."""
    n_rows = data.draw(st.integers(1, 5))
    table = [("p%d" % i, data.draw(st.integers(1, 8)), data.draw(st.lists(st.integers(0, 7), min_size=1, max_size=5, unique=True)), i) for i in range(n_rows)]
    rows = E.make_profiles(table)
    G = data.draw(st.integers(1, 12))
    occ = np.array(data.draw(st.lists(st.integers(0, 255), min_size=G, max_size=G)), dtype=np.uint8)
    quirks = data.draw(st.sampled_from([0, 3]))
    prof = np.array(data.draw(st.lists(st.integers(0, n_rows - 1), min_size=1, max_size=40)), dtype=np.uint8)
    node_off = np.array([0, G], dtype=np.uint32)
    fast = oracle.Fast(node_off, rows, quirks)
    fast.load(occ)
    cur = occ.copy()
    for p in prof:
        res = fast.place(alloc_requests(np.array([p], dtype=np.uint8)))[0]
        feasible = [g for g in range(G) if oracle.start_for(rows[p], quirks, int(cur[g])) != 9]
        if res["status"] == E.ST_PLACED:
            g, s, z = int(res["gpu"]), int(res["start"]), int(res["size"])
            assert g == feasible[0] and s == oracle.start_for(rows[p], quirks, int(cur[g])) and z == table[p][1]
            span = ((1 << z) - 1) << s
            assert cur[g] & span == 0                  # never double-books
            cur[g] |= span & 0xFF
        else:
            assert not feasible
        assert np.array_equal(fast.occupancy(), cur)
