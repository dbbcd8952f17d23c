"""Writes BASELINE configs 3 and 4 (first churn batches) as flat binaries for the CPU research prototypes in this directory."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import oracle
from instaslice_b200 import workloads as W
d = os.environ.get("ISL_PROTO_DIR", "/tmp/isl_proto")
os.makedirs(d, exist_ok=True)
ch = W.Churn()
fast = oracle.Fast(ch.node_off, ch.rows)
fast.load(np.zeros(ch.G, dtype=np.uint8))
st = {}
ch.generate(fast.place, after_prefill=lambda: st.update(occ0=fast.occupancy()))
batches = ch.batches[ch.n_prefill_batches:]
st["occ0"].tofile(f"{d}/c4_occ0.bin")
np.array([len(b) for b in batches], dtype=np.uint32).tofile(f"{d}/c4_sizes.bin")
np.concatenate(batches).tofile(f"{d}/c4_req.bin")
ch.rows.tofile(f"{d}/rows.bin")
node_off, occ, rows, req = W.config3()
occ.tofile(f"{d}/c3_occ0.bin"); req.tofile(f"{d}/c3_req.bin")
np.array([len(req)], dtype=np.uint32).tofile(f"{d}/c3_sizes.bin")
print("wrote", d)
