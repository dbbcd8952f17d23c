#!/usr/bin/env python
"""bench.py — MIG placements/s on BASELINE config 4 (1M ops x 65 536 GPUs, 50/50 alloc/free churn).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over the whole recorded workload: the inventory is reset to the
pre-filled state (64 KiB device copy, inside the timed region) and the 16 recorded batches (15 x 65 536 +
1 x 16 960 operations) are resolved in order.  The recording itself (which allocation each FREE names depends
on earlier placements) is made once, untimed, by running the generator through the engine.

value   whole-job placements/s with requests and results resident in HBM (isl_place_stream_device: ONE call per step, the
        engine pipelines the 16 batches over inventory segments inside one cooperative kernel)
e2e     the same through isl_place_stream with pinned HOST buffers: the H2D copy of every batch and the delivery of every
        result record into the caller's host array are inside the timed region (batches are fed on a second stream while the
        pipeline runs; an extra CTA writes finished chunks into the pinned result array)
N > 1   the inventory is partitioned over the ranks (contiguous GPU ranges); every rank holds the request stream and runs the
        segment pipeline over its own range; the per-profile queue-head token of every chunk crosses ranks INSIDE the running
        kernels (peer store into the next rank's inbox over NVLink, CUDA IPC); results are combined with an NCCL
        all-reduce(MIN) over the 8-byte records and the occupancy shards are all-gathered.  Strong scaling (the job is fixed).

The CPU oracle is used only for the cpu_baseline leg (timed baseline + parity check of the same sample) and
for --impl reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "placements_per_sec"
UNIT = "placements/s"
WORKLOAD = "C4: 8192 nodes x 8 GPUs (65536), H100-80GB table, prefill 50%, 1e6 ops, 50/50 alloc/free, batches of 65536, seed 42"


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one has been summarised."""
    try:
        with open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")) as f:
            return json.load(f).get("dram_bytes_per_launch")
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------------
def record_workload(E, W, torch):
    """Generate config 4 through the engine (untimed).  Returns (churn object, prefilled occupancy, churn batches, results)."""
    ch = W.Churn()
    eng = E.Engine(max_gpus=ch.G, max_batch=65536)
    eng.load_profiles(ch.rows)
    eng.load_inventory(ch.node_off, np.zeros(ch.G, dtype=np.uint8))
    results, snap = [], {}

    def placer(req):
        res = eng.place_batch(req)
        results.append(res)
        return res

    ch.generate(placer, after_prefill=lambda: snap.update(occ=eng.read_occupancy()))
    eng.close()
    nb = ch.n_prefill_batches
    return ch, snap["occ"], ch.batches[nb:], results[nb:]


def run_own(args):
    import torch
    import torch.distributed as dist
    from instaslice_b200 import engine as E
    from instaslice_b200 import workloads as W

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    os.environ["NCCL_DEBUG"] = os.environ.get("ISL_NCCL_DEBUG", "WARN")   # keep NCCL's version banner off stdout: rank 0 prints ONE JSON line
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()

    ch, occ0, batches, want = record_workload(E, W, torch)
    n_ops = sum(len(b) for b in batches)
    G = ch.G

    # engine on torch's current stream so that torch.cuda.Event brackets its kernels
    stream = torch.cuda.Stream()          # a real (non-default) stream: the legacy default stream has handle 0
    torch.cuda.set_stream(stream)
    eng = E.Engine(max_gpus=G, max_batch=1 << 20)
    eng.set_stream(stream.cuda_stream)
    eng.load_profiles(ch.rows)
    eng.load_inventory(ch.node_off, occ0)
    d_occ0 = torch.from_numpy(occ0).cuda()
    occ_view = _device_view(torch, eng.device_occupancy(), G)
    sizes = np.array([len(b) for b in batches], dtype=np.uint32)
    all_req = np.concatenate(batches).view(np.int64)
    d_in_all = torch.from_numpy(all_req.copy()).cuda()            # the whole stream, batch after batch, resident in HBM
    d_out_all = torch.empty_like(d_in_all)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    d_in = [d_in_all[offs[i]:offs[i + 1]] for i in range(len(batches))]
    d_out = [d_out_all[offs[i]:offs[i + 1]] for i in range(len(batches))]
    h_in_all = torch.from_numpy(all_req.copy()).pin_memory()
    h_out_all = torch.empty_like(h_in_all).pin_memory()
    h_out = [h_out_all[offs[i]:offs[i + 1]] for i in range(len(batches))]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2

    if world > 1:
        from instaslice_b200 import dist as D
        lo, hi = D.partition_bounds(G, world, rank)
        eng.set_partition(lo, hi)
        D.connect_ring(eng, rank, world)          # next rank's token inbox mapped through CUDA IPC (peer store over NVLink)
        stream_ids = iter(range(1, 1 << 30))

    def step_device():
        occ_view.copy_(d_occ0)
        if world == 1:      # ONE call for the stream of batches: the engine pipelines them over inventory segments
            eng.place_stream_ptr(sizes, d_in_all.data_ptr(), d_out_all.data_ptr(), device=True)
            return
        # every rank runs the segment pipeline over its own GPU range; tokens cross ranks inside the running kernels
        eng.place_stream_partitioned(sizes, d_in_all.data_ptr(), d_out_all.data_ptr(), next(stream_ids))
        D.merge_results(d_out_all)                                        # NCCL all-reduce(MIN) of the 8-byte records
        occ_view.copy_(D.gather_occupancy(occ_view[lo:hi], G, world, rank))   # NCCL all-gather of the occupancy shards

    def step_e2e():         # pinned host buffers in, pinned host buffers out: H2D + kernels + D2H inside the timed region
        occ_view.copy_(d_occ0)
        if world == 1:
            eng.place_stream_ptr(sizes, h_in_all.data_ptr(), h_out_all.data_ptr(), device=False)
            return
        d_in_all.copy_(h_in_all, non_blocking=True)                       # every rank stages the request stream
        eng.place_stream_partitioned(sizes, d_in_all.data_ptr(), d_out_all.data_ptr(), next(stream_ids))
        D.merge_results(d_out_all)
        occ_view.copy_(D.gather_occupancy(occ_view[lo:hi], G, world, rank))
        h_out_all.copy_(d_out_all, non_blocking=True)                     # merged results back to the host

    def timed(step_fn, steps, warmup, flush_l2=True):
        for _ in range(warmup):
            step_fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        total_ms = 0.0
        for _ in range(steps):
            if flush_l2:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step_fn()
            e1.record()
            e1.synchronize()
            total_ms += e0.elapsed_time(e1)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
        return total_ms

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.stats()["kernel_launches"]
    ms_dev = timed(step_device, args.steps, args.warmup)
    launches = eng.stats()["kernel_launches"] - launches0
    clocks = sampler.stop() if rank == 0 else None
    # parity gate on the last timed step: byte-identical to the recorded (single-GPU) results
    got = [t.cpu().numpy().view(E.RESULT_DTYPE) for t in d_out]
    parity = all(np.array_equal(a, b) for a, b in zip(got, want))

    e2e = None
    if world > 1:
        ms_e2e = timed(step_e2e, args.steps, args.warmup)
        parity = parity and all(np.array_equal(t.numpy().view(E.RESULT_DTYPE), b) for t, b in zip(h_out, want))
        e2e = {"value": n_ops * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": 8 * n_ops * world, "d2h_bytes_per_step": 8 * n_ops * world,
               "ms_per_step": ms_e2e / args.steps, "api": "isl_place_stream_partitioned per rank + NCCL merge; every rank copies the stream in and the merged results out"}
    if world == 1:
        ms_e2e = timed(step_e2e, args.steps, args.warmup)
        parity = parity and all(np.array_equal(t.numpy().view(E.RESULT_DTYPE), b) for t, b in zip(h_out, want))
        # the same job submitted batch by batch (synchronous per-batch calls, no cross-batch pipelining), for reference
        def step_per_batch():
            occ_view.copy_(d_occ0)
            for i, b in enumerate(batches):
                eng.place_batch_ptr(len(b), h_in_all.data_ptr() + 8 * int(offs[i]), h_out_all.data_ptr() + 8 * int(offs[i]))
        ms_pb = timed(step_per_batch, args.steps, 1)
        e2e = {"value": n_ops * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": 8 * n_ops, "d2h_bytes_per_step": 8 * n_ops,
               "ms_per_step": ms_e2e / args.steps, "api": "isl_place_stream (16 batches per call; batches fed and results delivered while the pipeline runs)",
               "per_batch_calls_value": n_ops * args.steps / (ms_pb / 1e3), "per_batch_calls_note": "isl_place_batch once per batch, synchronous"}

    roofline = cpu = None
    if rank == 0 and world == 1:
        # dominant kernel (k_chain), timed live with CUDA events on the engine's own stream in timing mode
        teng = E.Engine(max_gpus=G, max_batch=1 << 20, timing=True)
        teng.load_profiles(ch.rows)
        for rep in range(3):
            teng.load_inventory(ch.node_off, occ0)
            teng.reset_stats()
            teng.place_stream_ptr(sizes, d_in_all.data_ptr(), d_out_all.data_ptr(), device=True)
            teng.synchronize()
        st = teng.stats()
        ms_pipe = st["ms_commit"]                              # the single k_pipeline launch of the stream (CUDA events on the engine's stream)
        alg_bytes = 16 * n_ops + 2 * G * len(batches)         # B_alg = 16 R + 2 G per batch, whole stream = one launch
        peak, how = measured_peak()
        achieved = alg_bytes / (ms_pipe / 1e3) / 1e9
        # the sequential single-chain path on the same job, for the record (one k_chain launch per batch)
        seng = E.Engine(max_gpus=G, max_batch=1 << 20, timing=True, flags=E.FLAG_NO_PIPELINE)
        seng.load_profiles(ch.rows)
        seng.load_inventory(ch.node_off, occ0)
        seng.place_stream_ptr(sizes, d_in_all.data_ptr(), d_out_all.data_ptr(), device=True)
        seng.synchronize()
        sst = seng.stats()
        seng.close()
        roofline = {"bound": "hbm", "kernel": "k_pipeline", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(), "peak_source": how, "algorithmic_bytes_per_launch": alg_bytes,
                    "avg_launch_ms": ms_pipe, "launches": 1,
                    "note": "latency-bound exact commit chain pipelined over ~148 inventory segments (one CTA per SM); inventory, queues and candidates "
                            "are shared-memory / L2 resident by construction, so DRAM traffic stays below the algorithmic bytes",
                    "phase_ms_per_step": {"prepare": st["ms_free"], "partition": st["ms_partition"], "pipeline": st["ms_commit"], "total": st["ms_total"]},
                    "single_chain_path_ms_per_step": {"prepare": sst["ms_free"], "partition": sst["ms_partition"], "sweep": sst["ms_sweep"],
                                                      "chain+commit": sst["ms_commit"], "total": sst["ms_total"]}}
        teng.close()
        cpu = cpu_baseline(ch, occ0, batches, want)

    if rank == 0:
        line = {"metric": METRIC, "value": n_ops * args.steps / (ms_dev / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": WORKLOAD, "ops_per_step": n_ops, "alloc_requests_per_step": int(sum(int((b["op"] == E.OP_ALLOC).sum()) for b in batches)),
                           "ops_counted": "every request resolved: ALLOC decisions (placed or no-capacity) and FREEs", "batches_per_step": len(batches), "gpus_in_inventory": G,
                           "policy": "first-fit", "quirks": "REF_EXACT",
                           "parallelism": "segment pipeline, 1 GPU" if world == 1 else "inventory partitioned over %d ranks: peer-memory token ring + NCCL all-reduce(MIN) of results + NCCL all-gather of occupancy" % world,
                           "l2": "flushed between timed steps (256 MiB write)", "timing": "cuda events per step, max over ranks"},
                "parity": "bit-exact vs recorded single-GPU results" if parity else "MISMATCH",
                "gpu_launches": int(launches), "clocks": clocks}
        if e2e:
            line["e2e"] = e2e
        if roofline:
            line["roofline"] = roofline
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0 if parity else 1


def _device_view(torch, ptr: int, n: int):
    """uint8 torch view of engine-owned device memory (no copy)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 3}
    return torch.as_tensor(h, device="cuda")


# ---- CPU legs (the only places that touch oracle/) ------------------------------------------------------------
def _prefilled_faithful(oracle, ch, batches_prefill_results):
    f = oracle.Faithful(ch.node_off, ch.rows)
    pod = 0
    for req, res in batches_prefill_results:
        placed = (req["op"] == 0) & (res["status"] == 0)
        for g, s, z in zip(res["gpu"][placed].tolist(), res["start"][placed].tolist(), res["size"][placed].tolist()):
            f.add_allocation(g, s, z, pod)
            pod += 1
    return f


def cpu_baseline(ch, occ0, batches, want, sample_ops=1500):
    """ref_faithful (the reference as written, 1 reconcile worker) on the first `sample_ops` operations of churn
    batch 0 against the pre-filled inventory; ref_fast on the whole job.  Also the parity check of that sample."""
    import oracle
    from instaslice_b200 import engine as E
    f = oracle.Faithful(ch.node_off, ch.rows)
    f.load_occupancy_as_dangling(occ0)       # the pre-fill as realised slices
    req = batches[0][:sample_ops].copy()
    req = req[req["op"] == E.OP_ALLOC]        # frees of the sample would name pod-keyed entries; the sample times allocations
    t0 = time.perf_counter()
    res = f.place(req)
    dt = time.perf_counter() - t0
    fast = oracle.Fast(ch.node_off, ch.rows)
    fast.load(occ0)
    t1 = time.perf_counter()
    ok = True
    for b, w in zip(batches, want):
        ok = ok and np.array_equal(fast.place(b), w)
    dt_fast = time.perf_counter() - t1
    fast2 = oracle.Fast(ch.node_off, ch.rows)
    fast2.load(occ0)
    ok_sample = np.array_equal(fast2.place(req), res)
    return {"value": len(req) / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "ref_faithful.cpp on the %d ALLOC requests among the first %d ops of churn batch 0, pre-filled 65536-GPU inventory, %.1f s" % (len(req), sample_ops, dt),
            "ref_fast_value": sum(len(b) for b in batches) / dt_fast, "ref_fast_note": "bitmask restatement, whole job, 1 core (strong CPU baseline)",
            "parity_full_job_vs_ref_fast": bool(ok), "parity_sample_faithful_vs_fast": bool(ok_sample)}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (ref_faithful.cpp, the port — the Go binary cannot be built here)."""
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    import oracle
    from instaslice_b200 import engine as E
    from instaslice_b200 import workloads as W
    oracle.build()
    # the recorded workload needs a placer; the reference arm may execute oracle/, so ref_fast records it
    ch = W.Churn()
    fast = oracle.Fast(ch.node_off, ch.rows)
    fast.load(np.zeros(ch.G, dtype=np.uint8))
    state = {"occ0": None}
    ch.generate(fast.place, after_prefill=lambda: state.update(occ0=fast.occupancy()))
    batches = ch.batches[ch.n_prefill_batches:]
    sample_ops = args.sample
    req = batches[0][:sample_ops].copy()
    req = req[req["op"] == E.OP_ALLOC]
    times = []
    for i in range(args.warmup + args.steps):
        f = oracle.Faithful(ch.node_off, ch.rows)
        f.load_occupancy_as_dangling(state["occ0"])
        t0 = time.perf_counter()
        f.place(req)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = len(req) * args.steps / total
    sample = "ref_faithful.cpp, %d ALLOC requests among the first %d ops of churn batch 0 per step, pre-filled 65536-GPU inventory" % (len(req), sample_ops)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                      "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
                      "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                                       "note": "single reconcile worker like the reference (controller-runtime default); the Go binary cannot be built in this image; "
                                               "counts ALLOC decisions only (a FREE is a map delete by the daemonset in the reference), the GPU arm's ops are ~50% FREEs"},
                      "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--sample", type=int, default=1500, help="ops per step of the CPU reference arm")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "own" else args.warmup
    sys.exit(run_reference(args) if args.impl == "reference" else run_own(args))


if __name__ == "__main__":
    main()
