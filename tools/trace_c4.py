"""Timeline of the segment pipeline on config 4 (FLAG_TRACE): where does a stream step spend its time?"""
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
    eng.place_stream(batches)
tr = eng.read_trace().astype(np.int64)           # [chunk][seg][sweep done, token in, chain done, commit done]
t0 = tr[tr > 0].min()
tr = np.where(tr > 0, tr - t0, -1) / 1e3         # us
nc, ns, _ = tr.shape
print("stats", {k: v for k, v in eng.stats().items() if k.startswith("ms") or k.startswith("chain")})
print("kernel span us:", tr.max())
wait = tr[:, :, 1] - tr[:, :, 0]; chain = tr[:, :, 2] - tr[:, :, 1]; commit = tr[:, :, 3] - tr[:, :, 2]
print("per-chunk latency through all segments (token in seg0 -> chain done last seg), us:", np.round(tr[:, -1, 2] - tr[:, 0, 1], 1))
print("chunk c enters seg 0 at us:", np.round(tr[:, 0, 1], 1))
print("chunk c leaves last seg at us:", np.round(tr[:, -1, 2], 1))
busy = chain.sum(axis=0)
print("chain+window time per segment summed over chunks (us), top 12:", np.round(np.sort(busy)[-12:], 1), "argmax seg", int(busy.argmax()))
print("mean hop latency of idle segments (chain us where < 5):", float(chain[chain < 5].mean()), "count", int((chain < 5).sum()))
print("sum over segs of chain for chunk 8:", float(chain[8].sum()), " commit:", float(commit[8].sum()), " wait:", float(wait[8].sum()))
for s in (0, 1, 2, 8, 16, 32, 48, 63, 64, 100, 127):
    print("seg", s, "chunk8: wait %.1f chain %.1f commit %.1f | service per chunk (commit done c -> commit done c+1) mean %.1f" %
          (wait[8, s], chain[8, s], commit[8, s], float(np.diff(tr[:, s, 3]).mean())))
