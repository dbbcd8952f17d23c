"""Profile driver: BASELINE config 4 as ONE device-resident stream call (the launch bench.py's `replay` / causal-window modes make),
three times.  The workload is recorded through the single-chain path, so every k_pipeline launch of this process is the stream kernel:

    ncu --set full --clock-control none --import-source on -k regex:k_pipeline -s 2 -c 1 -o gpurun_out/r02_pipeline python tools/profile_c4_stream.py [window]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import __graft_entry__ as g
g.build()
from instaslice_b200 import engine as E, workloads as W

window = int(sys.argv[1]) if len(sys.argv) > 1 else 0
ch = W.Churn(min_age=max(1, window))
rec = E.Engine(max_gpus=ch.G, max_batch=65536, flags=E.FLAG_NO_PIPELINE)
rec.load_profiles(ch.rows); rec.load_inventory(ch.node_off, np.zeros(ch.G, dtype=np.uint8))
snap = {}
ch.generate(rec.place_batch, after_prefill=lambda: snap.update(occ=rec.read_occupancy()))
rec.close()
batches = ch.batches[ch.n_prefill_batches:]
sizes = np.array([len(b) for b in batches], dtype=np.uint32)
d_in = torch.from_numpy(np.concatenate(batches).view(np.int64).copy()).cuda()
d_out = torch.empty_like(d_in)
eng = E.Engine(max_gpus=ch.G, max_batch=1 << 20)
eng.load_profiles(ch.rows)
eng.set_causal_window(window)
for rep in range(3):
    eng.load_inventory(ch.node_off, snap["occ"])
    eng.place_stream_ptr(sizes, d_in.data_ptr(), d_out.data_ptr(), device=True)
    eng.synchronize()
print("stats", eng.stats())
