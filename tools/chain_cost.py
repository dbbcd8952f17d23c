"""Cost model of the segment pipeline's decision chain on config 4 (FLAG_TRACE).

For every (chunk, segment) cell the trace holds [sweep done, token in, chain done, commit done] globaltimer stamps; the
number of decisions of the cell is recounted from the results (PLACED records whose GPU lies in the segment).  A least
squares fit  chain_us = a + b * decisions  separates the per-decision latency of the loop (b) from the fixed cost of a
busy cell (a: heads, window staging, token publication).
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build()
from instaslice_b200 import engine as E, workloads as W

ch = W.Churn()
rec = E.Engine(max_gpus=ch.G, max_batch=65536)
rec.load_profiles(ch.rows); rec.load_inventory(ch.node_off, np.zeros(ch.G, dtype=np.uint8))
snap = {}
ch.generate(rec.place_batch, after_prefill=lambda: snap.update(occ=rec.read_occupancy()))
batches = ch.batches[ch.n_prefill_batches:]
eng = E.Engine(max_gpus=ch.G, max_batch=1 << 20, flags=E.FLAG_TRACE, timing=True)
eng.load_profiles(ch.rows)
for rep in range(3):
    eng.load_inventory(ch.node_off, snap["occ"]); eng.reset_stats()
    res = eng.place_stream(batches)
tr = eng.read_trace().astype(np.int64)
chk = E.Engine(max_gpus=ch.G, max_batch=1 << 20, flags=E.FLAG_NO_PIPELINE)      # the single-chain path as cross-check
chk.load_profiles(ch.rows); chk.load_inventory(ch.node_off, snap["occ"])
same = all(np.array_equal(chk.place_batch(b), r) for b, r in zip(batches, res)) and np.array_equal(chk.read_occupancy(), eng.read_occupancy())
print("pipeline == single-chain path:", same)
nc, ns, _ = tr.shape
seg = -(-ch.G // ns // 64) * 64 if ns * 64 < ch.G else 64
while seg * ns < ch.G: seg += 64
dec = np.zeros((nc, ns), dtype=np.int64)
assert nc == len(batches), (nc, len(batches))
for c, r in enumerate(res):
    placed = r[(r["status"] == E.ST_PLACED) & (batches[c]["op"] == E.OP_ALLOC)]
    dec[c] = np.bincount(placed["gpu"] // seg, minlength=ns)[:ns]
if os.environ.get("ISL_HACK"):      # timing-only builds (wrong results on purpose): take the decision counts from the trace
    dec = tr[:, :, 6].copy()
chain = (tr[:, :, 2] - tr[:, :, 1]) / 1e3
commit = (tr[:, :, 3] - tr[:, :, 2]) / 1e3
busy = dec > 0
A = np.stack([np.ones(busy.sum()), dec[busy]], axis=1)
(a, b), *_ = np.linalg.lstsq(A, chain[busy], rcond=None)
print("segments", ns, "x", seg, "GPUs; busy cells", int(busy.sum()), "decisions", int(dec.sum()))
print("chain_us = %.2f + %.4f * decisions   (%.1f ns per decision; %.0f cycles at 1.965 GHz)" % (a, b, b * 1e3, b * 1965))
(a2, b2), *_ = np.linalg.lstsq(A, commit[busy], rcond=None)
print("commit_us = %.2f + %.4f * decisions" % (a2, b2))
print("idle cell hop (token in -> token out) us: mean %.2f" % float(chain[~busy].mean()))
print("kernel span us %.1f; sum of busy-cell chain for chunk 0: %.1f; chunk 8: %.1f" % ((tr[:, :, :4].max() - tr[:, :, :4][tr[:, :, :4] > 0].min()) / 1e3, chain[0][busy[0]].sum(), chain[8][busy[8]].sum()))
print("per-chunk: decisions", dec.sum(axis=1).tolist())
print("per-chunk busy segments", busy.sum(axis=1).tolist())
# finer split of a busy cell (trace words 4..7)
t_in, t_out, t_cs, t_ce = tr[:, :, 1], tr[:, :, 2], tr[:, :, 4], tr[:, :, 5]
ndec, jv = tr[:, :, 6], tr[:, :, 7]
jumps, visited = jv & 0xFFFFFFFF, jv >> 32
assert (ndec[busy] == dec[busy]).all(), "trace decision counts disagree with the results"
pre = (t_cs - t_in)[busy] / 1e3; loop = (t_ce - t_cs)[busy] / 1e3; post = (t_out - t_ce)[busy] / 1e3
print("busy cell: token in -> loop start %.2f us | loop %.2f us | loop end -> token out %.2f us (means)" % (pre.mean(), loop.mean(), post.mean()))
t_h, t_st = tr[:, :, 8], tr[:, :, 9]
print("  token in -> heads %.2f us -> windows staged %.2f us -> loop start %.2f us" % (((t_h - t_in)[busy] / 1e3).mean(), ((t_st - t_h)[busy] / 1e3).mean(), ((t_cs - t_st)[busy] / 1e3).mean()))
wt = tr[:, :, 11]
print("  warp 0 done converting its window %.2f us after heads (window of the first profile: mean %.0f entries; free usable slices on candidates: mean %.0f)" %
      (((tr[:, :, 10] - t_h)[busy] / 1e3).mean(), (wt & 0xFFFFFFFF)[busy].mean(), (wt >> 32)[busy].mean()))
B = np.stack([np.ones(busy.sum()), dec[busy], jumps[busy], visited[busy]], axis=1)
coef, *_ = np.linalg.lstsq(B, loop, rcond=None)
print("loop_us = %.3f + %.4f*decisions + %.4f*jumps + %.4f*visited" % tuple(coef))
print("means per busy cell: decisions %.0f jumps %.1f visited %.0f" % (dec[busy].mean(), jumps[busy].mean(), visited[busy].mean()))
