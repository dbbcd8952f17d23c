"""BASELINE config 5: vLLM-shaped replay (samples/vllm_dep.yaml requests nvidia.com/mig-3g.20gb) — placement latency.

Open-loop Poisson arrivals at RATE req/s for DURATION s, 100 % 3g.20gb on A100-40GB tables, 4096 GPUs, slice lifetimes
exponential with mean 30 s.  The replay runs in real time against the wall clock: the driver loop
hands the engine every request (and every expired slice as a FREE) that has arrived since the previous call, through
isl_place_batch with host buffers, exactly as a reconciler would.  Latency = result available - arrival.
Prints one JSON line.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build()
from instaslice_b200 import engine as E, tables, workloads as W

RATE = float(os.environ.get("C5_RATE", 10000)); DURATION = float(os.environ.get("C5_SECONDS", 5)); MEAN_LIFE = 30.0
rng = W.SplitMix64(42)
n = int(RATE * DURATION)
u = (rng.next(n) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
arrivals = np.cumsum(-np.log1p(-u) / RATE)
life = -np.log1p(-(rng.next(n) >> np.uint64(11)).astype(np.float64) / float(1 << 53)) * MEAN_LIFE
prof = tables.profile_index(tables.A100_40GB, "3g.20gb")
G = 4096
eng = E.Engine(max_gpus=G, max_batch=65536, flags=int(os.environ.get("C5_FLAGS", "0")))
eng.load_profiles(E.make_profiles(tables.A100_40GB))
eng.load_inventory(W.node_offsets(G // 8, 8), np.zeros(G, dtype=np.uint8))
for _ in range(200):                                   # warm the path (kernels loaded, buffers allocated)
    eng.place_batch(np.array([(0, E.PROFILE_UNKNOWN, E.OP_NOOP, 0, 0)], dtype=E.REQUEST_DTYPE))
import heapq
expiry = []                                            # (time, gpu, start, size)
lat = np.zeros(n); placed = 0; calls = 0; batch_sizes = []
i = 0
t0 = time.perf_counter()
while i < n:
    now = time.perf_counter() - t0
    j = i
    while j < n and arrivals[j] <= now:
        j += 1
    frees = []
    while expiry and expiry[0][0] <= now:
        frees.append(heapq.heappop(expiry))
    if j == i and not frees:
        continue
    req = np.zeros((j - i) + len(frees), dtype=E.REQUEST_DTYPE)
    for k, (_, gpu, s, z) in enumerate(frees):
        req[k] = (gpu, 0, E.OP_FREE, s, z)
    req["profile"][len(frees):] = prof
    res = eng.place_batch(req)
    done = time.perf_counter() - t0
    calls += 1; batch_sizes.append(len(req))
    for k in range(i, j):
        lat[k] = done - arrivals[k]
        r = res[len(frees) + k - i]
        if r["status"] == E.ST_PLACED:
            placed += 1
            heapq.heappush(expiry, (arrivals[k] + life[k], int(r["gpu"]), int(r["start"]), int(r["size"])))
    i = j
wall = time.perf_counter() - t0
us = lat * 1e6
print(json.dumps({"config": "C5: Poisson %.0f req/s x %.0f s, 100%% 3g.20gb, A100-40GB tables, 4096 GPUs, exp(30 s) lifetimes" % (RATE, DURATION),
                  "requests": n, "placed": placed, "calls": calls, "mean_batch": float(np.mean(batch_sizes)), "wall_s": wall,
                  "latency_us": {"p50": float(np.percentile(us, 50)), "p90": float(np.percentile(us, 90)), "p99": float(np.percentile(us, 99)),
                                 "p999": float(np.percentile(us, 99.9)), "max": float(us.max()), "mean": float(us.mean())},
                  "api": "isl_place_batch, host buffers, one call per driver-loop turn"}))
