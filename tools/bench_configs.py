"""Per-config numbers for BASELINE configs 1-3 (config 4 is bench.py, config 5 is tools/latency_c5.py).

For each config: e2e placements/s through isl_place_batch with host buffers (best of 5), kernel-only time from the
engine's CUDA-event statistics, the two CPU baselines on the same input (ref_faithful = the reference as written,
bounded sample where the full run would take minutes; ref_fast = bitmask restatement) and the parity verdict
(results and final occupancy byte-identical to ref_fast).  One JSON line per config.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build()
import oracle
from instaslice_b200 import engine as E, workloads as W


def run(name, node_off, occ, rows, req, policy=E.POLICY_FIRST_FIT, faithful_sample=None):
    G = len(occ)
    eng = E.Engine(max_gpus=max(4096, G), max_batch=1 << 20, policy=policy, timing=True)
    eng.load_profiles(rows)
    best = 1e9
    for rep in range(5):
        eng.load_inventory(node_off, occ); eng.reset_stats()
        t0 = time.perf_counter(); got = eng.place_batch(req); dt = time.perf_counter() - t0
        best = min(best, dt)
    st = eng.stats()
    fast = oracle.Fast(node_off, rows, 3, policy=policy); fast.load(occ)
    t0 = time.perf_counter(); want = fast.place(req); t_fast = time.perf_counter() - t0
    parity = bool(np.array_equal(got, want) and np.array_equal(eng.read_occupancy(), fast.occupancy()))
    line = {"config": name, "requests": len(req), "gpus": G, "policy": "best-fit" if policy else "first-fit",
            "e2e_placements_per_s": len(req) / best, "e2e_ms": best * 1e3, "kernel_ms": st["ms_total"], "kernel_placements_per_s": len(req) / (st["ms_total"] / 1e3),
            "placed": int((got["status"] == E.ST_PLACED).sum()), "parity_vs_ref_fast": parity,
            "ref_fast_placements_per_s": len(req) / t_fast, "algorithmic_bytes": 16 * len(req) + 2 * G,
            "roofline_frac_of_measured_6575GBs": (16 * len(req) + 2 * G) / (st["ms_total"] / 1e3) / 6575.1e9}
    if policy == E.POLICY_FIRST_FIT:
        k = len(req) if faithful_sample is None else min(faithful_sample, len(req))
        f = oracle.Faithful(node_off, rows); f.load_occupancy_as_dangling(occ)
        t0 = time.perf_counter(); fr = f.place(req[:k]); t_f = time.perf_counter() - t0
        line.update({"ref_faithful_placements_per_s": k / t_f, "ref_faithful_sample": k, "ref_faithful_matches": bool(np.array_equal(fr, want[:k])),
                     "speedup_e2e_vs_ref_faithful": (len(req) / best) / (k / t_f), "speedup_e2e_vs_ref_fast": (len(req) / best) / (len(req) / t_fast)})
    print(json.dumps(line), flush=True)


run("C1 samples/test-pod.yaml: one 1g.5gb on 1 empty A100-40GB GPU", *W.config1())
run("C2 10k x 1g.10gb on 256 GPUs", *W.config2())
c3 = W.config3()
run("C3 100k mixed on 4096 GPUs", *c3, faithful_sample=3000)
run("C3 100k mixed on 4096 GPUs", *c3, policy=E.POLICY_BEST_FIT)
