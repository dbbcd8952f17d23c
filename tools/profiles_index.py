"""Regenerates the measured-numbers table of profiles/README.md from the recorded bench lines (profiles/r02_bench_*.json).

    python tools/profiles_index.py > profiles/r02_table.md
"""
import glob
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load(name):
    p = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(p):
        return None
    with open(p) as f:
        lines = [l for l in f.read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def fmt(v):
    if v is None:
        return "—"
    if v >= 1e6:
        return "%.1f M/s" % (v / 1e6)
    if v >= 1e3:
        return "%.1f k/s" % (v / 1e3)
    return "%.0f /s" % v


print("| config | `value` (device-resident) | `e2e` (host buffers through the C ABI) | `ref_fast` (bitmask CPU code, 1 core) | reference as written (CPU port, 1 worker) | parity | roofline frac (B_alg / t / 6575 GB/s) |")
print("|---|---|---|---|---|---|---|")
for c, label in (("c1", "C1 one 1g.5gb pod, 1 GPU"), ("c2", "C2 10k x 1g.10gb, 256 GPUs"), ("c3", "C3 100k mixed, 4096 GPUs, first-fit"),
                 ("c3bf", "C3 best-fit (extension)"), ("c4", "C4 10^6 ops, 65 536 GPUs, churn — causal feed")):
    d = load("r02_bench_%s.json" % c)
    if not d:
        continue
    cb = d.get("cpu_baseline", {})
    faithful = cb.get("value") if "parity_sample_faithful_vs_fast" in cb else None
    extra = ""
    if c == "c4":
        label = "C4 10^6 ops, 65 536 GPUs, churn — STRICT CAUSAL (headline): the original stream, ONE batch in flight, speculative rounds; ALLOC decisions" if d["config"]["min_age_batches"] == 1 else label
    print("| %s%s | %s (%.3g ms) | %s (%.3g ms) | %s | %s | %s | %.2e |" % (
        label, extra, fmt(d["value"]), d["ms_per_step"], fmt(d["e2e"]["value"]), d["e2e"]["ms_per_step"], fmt(cb.get("ref_fast_value")), fmt(faithful),
        "bit-exact" if d["parity"].startswith("bit-exact") else "MISMATCH", d["roofline"]["frac"] or 0))
    if c == "c4":
        print("| ... the same through the open-stream API (pinned host buffers) | | %s (%.3g ms) | | | bit-exact | |" % (fmt(d["e2e"]["open_stream_value"]), d["e2e"]["open_stream_ms_per_step"]))
        f = d["causal_feed"]
        print("| C4 causal feed (FREEs name allocations >= 2 batches old, TWO batches in flight) | %s (%.3g ms) | %s open stream | %s | | %s | |" % (
            fmt(f["value"]), f["ms_per_step"], fmt(f["open_stream_e2e_value"]), fmt(f["ref_fast_value"]), "bit-exact" if f["parity_vs_ref_fast"] else "MISMATCH"))
        print("| C4 replay (original stream handed over at once; NOT causally available) | %s (%.3g ms) | %s | | | bit-exact | |" % (
            fmt(d["replay_value"]), d["replay_ms_per_step"], fmt(d["replay_e2e_value"])))
        sp = d["roofline"].get("speculative_rounds")
        if sp:
            print("| (C4 headline: %.1f rounds and %.0f segment simulations per batch) | | | | | | |" % (sp["rounds_per_chunk"], sp["simulations_per_chunk"]))
d = load("r02_bench_c5.json")
if d:
    l, cb = d["latency_us"], d["cpu_baseline"]
    print()
    print("C5 (10 s open-loop replay at 10 000 req/s, 3g.20gb, 4096 GPUs, native driver): p50 %.1f us, p90 %.1f us, p99 %.1f us, max %.0f us per request "
          "(arrival -> result through `isl_place_batch`); the same trace with `ref_fast` as the placer: p50 %.2f us / p99 %.2f us; with the reference as written "
          "(first 2 s): p50 %.1f s (saturated: it resolves ~10^4 pods/s on 4096 GPUs).  %s." % (
              l["p50"], l["p90"], l["p99"], l["max"], cb["ref_fast_latency_us"]["p50"], cb["ref_fast_latency_us"]["p99"], cb["ref_faithful_latency_us"]["p50"] / 1e6,
              "Every call's results replayed through the oracle: bit-exact" if d["parity"].startswith("bit-exact") else "MISMATCH"))
rows = []
for n in (1, 2, 4, 8):
    d = load("r02_bench_c4.json") if n == 1 else load("r02_bench_n%d.json" % n)
    if d:
        ph = d["roofline"].get("phase_ms_per_step_max_over_ranks", {})
        rows.append((n, d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], ph, d["parity"].startswith("bit-exact")))
if rows:
    print()
    print("| N GPUs (config 4 strict causal, inventory partitioned over the ranks; strong scaling: the job is fixed) | ms per step (max over ranks) | ALLOC decisions/s | e2e | roofline frac | pre-pass + pipeline | occupancy all-gather | result merge | parity vs `ref_fast` |")
    print("|---|---|---|---|---|---|---|---|---|")
    for n, ms, v, e, ems, fr, ph, ok in rows:
        vals = list(ph.values())
        print("| %d | %.3f | %s | %s (%.3g ms) | %.2e | %s | %s | %s | %s |" % (n, ms, fmt(v), fmt(e), ems, fr, ("%.3f ms" % vals[0]) if vals else "—", ("%.3f ms" % vals[1]) if len(vals) > 1 else "—",
                                                                            "0 (peer stores inside the kernel)" if n > 1 else "—", "bit-exact" if ok else "MISMATCH"))
