"""Per-batch diagnostics of the C4 workload: chain steps / visited / jumps / phase times."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g
g.build()
from instaslice_b200 import engine as E, workloads as W

ch = W.Churn()
eng = E.Engine(max_gpus=ch.G, max_batch=65536, timing=True)
eng.load_profiles(ch.rows)
eng.load_inventory(ch.node_off, np.zeros(ch.G, dtype=np.uint8))
rows = []
def placer(req):
    eng.reset_stats()
    res = eng.place_batch(req)
    st = eng.stats()
    rows.append((len(req), int((req["op"] == 0).sum()), st["placed"], st["chain_steps"], st["chain_gpus_visited"], st["chain_jumps"],
                 round(st["ms_commit"], 3), round(st["ms_sweep"], 3), round(st["ms_partition"], 3), round(st["ms_free"], 3)))
    return res
ch.generate(placer)
print("n allocs placed steps visited jumps ms_commit ms_sweep ms_part ms_prep")
for i, r in enumerate(rows):
    if i >= ch.n_prefill_batches - 2:
        print(i, *r)
