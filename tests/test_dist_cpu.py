"""Host-side logic of the partitioned (N > 1) path on CPU: world_size-2 gloo processes.

The product kernels need a GPU; what runs here is (1) ``dist.partition_bounds`` / ``merge_results`` / ``gather_occupancy``
over real ``torch.distributed`` collectives and (2) the composition argument the partitioned path rests on: resolving
rank ranges in order, each starting from the queue-head token of the previous rank, equals the global sequential
first-fit.  The per-range resolver here is a small pure-Python GPU-major stream filter (test-only model of the chain),
checked against the request-major oracle.
"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from instaslice_b200 import dist as D
from instaslice_b200 import engine as E
from instaslice_b200 import tables, workloads as W


def test_partition_bounds_cover_and_align():
    for G in (1, 7, 512, 4096, 65536, 65537, 100000):
        for world in (1, 2, 3, 4, 8):
            b = D.all_bounds(G, world)
            assert b[0][0] == 0 and b[-1][1] == G
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            assert all(lo <= hi for lo, hi in b)
            if G >= world * D.SEGMENT:
                assert all(lo % D.SEGMENT == 0 for lo, _ in b)
    assert D.all_bounds(65536, 8) == [(i * 8192, (i + 1) * 8192) for i in range(8)]


def chain_range(occ, lo, hi, lut, rows, queues, heads, out):
    """GPU-major stream filtering of canonical GPUs [lo, hi): GPU g accepts, in request order, a prefix of each
    profile's remaining queue.  ``heads`` is the token; results go to ``out``; occupancy is updated in place."""
    for g in range(lo, hi):
        while True:
            best = None
            for p, q in enumerate(queues):
                if heads[p] < len(q) and lut[p][occ[g]] != 9:
                    if best is None or q[heads[p]] < queues[best][heads[best]]:
                        best = p
            if best is None:
                break
            s = int(lut[best][occ[g]])
            size = int(rows[best]["size"])
            occ[g] |= ((1 << size) - 1) << s
            out[queues[best][heads[best]]] = (g, s, size, E.ST_PLACED)
            heads[best] += 1


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        table = tables.H100_80GB
        rows = E.make_profiles(table)
        rng = W.SplitMix64(99)
        G = 1500
        node_off = W.node_offsets(G // 4, 4)
        occ0 = ((rng.next(G) & rng.next(G)) & np.uint64(0x7F)).astype(np.uint8)
        req = W.alloc_requests(W.mix_profiles(rng, 3000))
        lut = [[oracle.start_for(rows[p], 3, o) for o in range(256)] for p in range(len(rows))]
        queues = [[i for i in range(len(req)) if req["profile"][i] == p] for p in range(len(rows))]
        lo, hi = D.partition_bounds(G, world, rank, align=64)
        # token chain: rank d starts from the heads rank d-1 ended with
        heads = torch.zeros(16, dtype=torch.int64)
        if rank > 0:
            dist.recv(heads, src=rank - 1)
        h = heads.tolist()
        occ = occ0.copy()
        out = np.zeros(len(req), dtype=E.RESULT_DTYPE)
        out["gpu"], out["start"], out["status"] = E.GPU_NONE, 9, E.ST_NO_CAPACITY
        out["size"] = rows["size"][req["profile"]]
        chain_range(occ, lo, hi, lut, rows, queues, h, out)
        if rank < world - 1:
            dist.send(torch.tensor(h + [0] * (16 - len(h)), dtype=torch.int64)[:16], dst=rank + 1)
        merged = D.merge_results(torch.from_numpy(out.view(np.int64).copy())).numpy().view(E.RESULT_DTYPE)
        full = D.gather_occupancy(torch.from_numpy(occ[lo:hi].copy()), G, world, rank, align=64).numpy()
        ref = oracle.Fast(node_off, rows)
        ref.load(occ0)
        want = ref.place(req)
        assert np.array_equal(merged, want), "merged results differ from the global sequential first-fit"
        assert np.array_equal(full, ref.occupancy()), "gathered occupancy differs"
        open(os.path.join(tmp, "ok%d" % rank), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_token_chain_merge_and_gather_gloo(tmp_path, world):
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / ("ok%d" % r)).exists() for r in range(world))
