"""Timeline of the speculative rounds on config 4 (strict causal: one batch in flight) from the FLAG_TRACE stamps: per chunk the rounds,
the simulations per stage, when each stage was certified (us after the chunk's first stage started) — the frontier of exactness over time.
    python tools/spec_trace.py [chunk]
Per-round stamps of ONE (chunk, stage) cell (start / heads / chain / publish / gather / correction) need a debugging build — the stamps
next to the decision loop cost ~14 % of a round and are compiled out by default:
    tools/ab_build.sh dbg "-DISL_SPEC_DBG_STAMPS";  ISL_LIB=$PWD/tools/_ab/dbg.so ISL_SPEC_DBG=2,40 python tools/spec_trace.py 2"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build()
from instaslice_b200 import engine as E, workloads as W
import oracle

ch = W.Churn()
fast = oracle.Fast(ch.node_off, ch.rows)
fast.load(np.zeros(ch.G, dtype=np.uint8))
snap = {}
ch.generate(fast.place, after_prefill=lambda: snap.update(occ=fast.occupancy()))
batches = ch.batches[ch.n_prefill_batches:]
eng = E.Engine(max_gpus=ch.G, max_batch=1 << 20, flags=E.FLAG_TRACE)
eng.load_profiles(ch.rows)
eng.set_causal_window(int(os.environ.get("WINDOW", "1")))
for rep in range(3):
    eng.load_inventory(ch.node_off, snap["occ"]); eng.reset_stats()
    res = eng.place_stream(batches)
st = eng.stats()
print("stats: spec_chunks %d rounds %d (%.1f per chunk) sims %d (%.1f per chunk)" % (st["spec_chunks"], st["spec_rounds"], st["spec_rounds"] / max(1, st["spec_chunks"]), st["spec_sims"], st["spec_sims"] / max(1, st["spec_chunks"])))
tr = eng.read_trace().astype(np.int64)
nc, ns, _ = tr.shape
print("chunks", nc, "stages", ns)
t0 = tr[:, :, 0]; cert = tr[:, :, 2]; commit = tr[:, :, 3]; nlog = tr[:, :, 6]; sims = tr[:, :, 7]; rounds = tr[:, :, 11]
for c in range(nc):
    base = t0[c][t0[c] > 0].min()
    print("chunk %2d: start->all committed %.1f us | rounds (last stage) %d | max rounds %d | decisions %d | busy stages %d | sims total %d" % (
        c, (commit[c].max() - base) / 1e3, rounds[c, -1], rounds[c].max(), nlog[c].sum(), int((nlog[c] > 0).sum()), sims[c].sum()))
if nc > 1:
    print("chunk-to-chunk period us:", np.round(np.diff(commit.max(axis=1)) / 1e3, 1).tolist())
c = int(sys.argv[1]) if len(sys.argv) > 1 else min(2, nc - 1)
base = t0[c][t0[c] > 0].min()
print("chunk %d per stage: [stage] prediction ready us, certified us, round, sims, decisions" % c)
for s_ in range(ns):
    if s_ % 4 == 0 or s_ == ns - 1:
        print("  [%3d] %7.1f %7.1f  r%-3d s%-3d d%-4d" % (s_, (t0[c, s_] - base) / 1e3, (cert[c, s_] - base) / 1e3, rounds[c, s_], sims[c, s_], nlog[c, s_]))

if os.environ.get("ISL_SPEC_DBG"):
    import ctypes as C
    lib = E.load_library()
    buf = np.zeros(160 * 8, dtype=np.uint64)
    lib.isl_debug_spec_rounds.restype = C.c_int
    rc = lib.isl_debug_spec_rounds(eng._h, buf.ctypes.data_as(C.c_void_p), C.c_uint32(buf.size))
    d = buf.reshape(160, 8).astype(np.int64)
    print("per-round stamps of cell", os.environ["ISL_SPEC_DBG"], "rc", rc, "(us since the cell's first round): start | heads | chain start | chain end | published | gathered | new prediction | decisions, simulated")
    b0 = d[1, 0]
    for r in range(1, 160):
        if d[r, 0] == 0: break
        rel = lambda x: (x - b0) / 1e3 if x else float("nan")
        print("  r%-3d %8.2f | %7.2f | %7.2f | %7.2f | %7.2f | %7.2f | %7.2f | d%d sim%d" % (r, rel(d[r, 0]), rel(d[r, 1]), rel(d[r, 2]), rel(d[r, 3]), rel(d[r, 4]), rel(d[r, 5]), rel(d[r, 6]), d[r, 7] & 0xFFFFFFFF, d[r, 7] >> 32))
