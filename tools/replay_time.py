"""Times config 4 as ONE device-resident stream call (plain pipeline: the replay ceiling) and with a causal window of 1 (speculative rounds).
    ISL_LIB=tools/_ab/<variant>.so python tools/replay_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from instaslice_b200 import engine as E, workloads as W
import oracle
ch = W.Churn()
fast = oracle.Fast(ch.node_off, ch.rows); fast.load(np.zeros(ch.G, dtype=np.uint8))
snap = {}
ch.generate(fast.place, after_prefill=lambda: snap.update(occ=fast.occupancy()))
batches = ch.batches[ch.n_prefill_batches:]
sizes = np.array([len(b) for b in batches], dtype=np.uint32)
d_in = torch.from_numpy(np.concatenate(batches).view(np.int64).copy()).cuda()
d_out = torch.empty_like(d_in)
eng = E.Engine(max_gpus=ch.G, max_batch=1 << 20)
stream = torch.cuda.Stream()
eng.set_stream(stream.cuda_stream)
eng.load_profiles(ch.rows)
for window in (0, 1):
    eng.set_causal_window(window)
    ts = []
    for rep in range(8):
        eng.load_inventory(ch.node_off, snap["occ"])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); eng.place_stream_ptr(sizes, d_in.data_ptr(), d_out.data_ptr(), device=True); e1.record(stream); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("window %d: ms per stream call (last 5 of 8): %s  min %.3f" % (window, " ".join("%.3f" % t for t in ts[3:]), min(ts[3:])))
