"""Multi-GPU plumbing for a partitioned inventory (BASELINE config 4).

One process per GPU (``torch.distributed``, NCCL on B200s, gloo in the CPU tests).  Rank *d* owns the contiguous
canonical GPU range ``partition_bounds(G, world, d)``; every rank holds the whole request stream.

* the queue-head token of every chunk crosses rank boundaries INSIDE the running segment-pipeline kernels through
  CUDA-IPC peer memory (``connect_ring``) — no host round trip, no collective on the data path;
* with ``connect_spec`` the stages of all ranks form ONE sequence that resolves a batch by speculative rounds: every
  rank maps every other rank's record memory and a stage stores the per-round records later ranks read straight into their copies;
* results are gathered on the OWNER rank (rank 0, where the controller runs) without a collective: ``connect_owner`` maps
  the owner's result array into every other rank, and the commit threads of those ranks store each PLACED record there
  as well (peer store inside the running kernel).  ``merge_results`` (all-reduce(MIN) over the 8-byte records: a PLACED
  record sorts below a NO_CAPACITY one) remains for callers that want the full result array on EVERY rank;
* occupancy: ``gather_occupancy`` all-gathers the per-rank shards (8 KiB per rank at config 4).

Between two partitioned stream calls all ranks must pass a collective (``merge_results`` is one): the inbox slots of
a stream may only be overwritten once the next rank has consumed them.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

SEGMENT = 512     # GPUs per pipeline segment at full size: rank boundaries are kept on segment boundaries


def partition_bounds(G: int, world: int, rank: int, align: int = SEGMENT) -> tuple[int, int]:
    """Contiguous, ascending, exhaustive ranges; interior boundaries are multiples of ``align`` when G allows it."""
    def cut(r):
        if r <= 0:
            return 0
        if r >= world:
            return G
        x = (G * r) // world
        if G >= world * align:
            x = (x + align // 2) // align * align
        return min(G, x)
    return cut(rank), cut(rank + 1)


def all_bounds(G: int, world: int, align: int = SEGMENT):
    return [partition_bounds(G, world, r, align) for r in range(world)]


def connect_ring(engine, rank: int, world: int, group=None):
    """Exchange the inbox IPC handles and map the next rank's inbox into this rank (peer store over NVLink)."""
    handles = [None] * world
    dist.all_gather_object(handles, engine.ipc_inbox_handle(), group=group)
    engine.ipc_connect(handles[rank + 1] if rank + 1 < world else None, has_prev=rank > 0)


def connect_owner(engine, rank: int, world: int, group=None):
    """Map the owner's (rank 0) result array into every other rank; tell every engine the size of the ring (causal window)."""
    handles = [None] * world
    dist.all_gather_object(handles, engine.ipc_results_handle() if rank == 0 else None, group=group)
    if rank > 0:
        engine.ipc_connect_owner(handles[0])
    engine.set_ring_world(world)


def connect_spec(engine, rank: int, world: int, G: int, group=None, align: int = SEGMENT):
    """Speculative rounds across ranks: every rank maps every other rank's record memory (the per-round records of the stages cross
    ranks as peer stores); ``partition_bounds`` must be what the engines' partitions are."""
    handles = [None] * world
    dist.all_gather_object(handles, engine.ipc_spec_handle(), group=group)
    bounds = [lo for lo, _ in all_bounds(G, world, align)] + [G]
    engine.ipc_connect_spec(world, rank, handles, bounds)


def merge_results(records_i64: torch.Tensor, group=None) -> torch.Tensor:
    """Element-wise MIN over ranks of the 8-byte result records viewed as int64 (in place)."""
    dist.all_reduce(records_i64, op=dist.ReduceOp.MIN, group=group)
    return records_i64


def gather_occupancy(shard: torch.Tensor, G: int, world: int, rank: int, group=None, align: int = SEGMENT) -> torch.Tensor:
    """All-gather the per-rank occupancy shards (uint8) into the full G-byte inventory, in canonical order."""
    bounds = all_bounds(G, world, align)
    width = max(hi - lo for lo, hi in bounds)
    padded = torch.zeros(width, dtype=torch.uint8, device=shard.device)
    padded[: shard.numel()] = shard
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, bounds)])
