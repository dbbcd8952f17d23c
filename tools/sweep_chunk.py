"""Stream step time of config 4 versus pipeline chunk size (ISL_PIPE_CHUNK)."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import numpy as np
    from instaslice_b200 import engine as E, workloads as W
    ch = W.Churn()
    rec = E.Engine(max_gpus=ch.G, max_batch=65536, flags=E.FLAG_NO_PIPELINE)
    rec.load_profiles(ch.rows); rec.load_inventory(ch.node_off, np.zeros(ch.G, dtype=np.uint8))
    snap = {}; res = []
    def placer(r):
        x = rec.place_batch(r); res.append(x); return x
    ch.generate(placer, after_prefill=lambda: snap.update(occ=rec.read_occupancy()))
    batches = ch.batches[ch.n_prefill_batches:]; want = res[ch.n_prefill_batches:]
    eng = E.Engine(max_gpus=ch.G, max_batch=1 << 20, timing=True)
    eng.load_profiles(ch.rows)
    best = 1e9
    for rep in range(5):
        eng.load_inventory(ch.node_off, snap["occ"]); eng.reset_stats()
        got = eng.place_stream(batches)
        st = eng.stats(); best = min(best, st["ms_total"])
    ok = all(np.array_equal(a, b) for a, b in zip(got, want))
    print(json.dumps({"chunk": os.environ.get("ISL_PIPE_CHUNK"), "segments": os.environ.get("ISL_PIPE_SEGMENTS"), "ms_total_best": round(best, 3), "ms_pipeline": round(st["ms_commit"], 3),
                      "ms_prepare": round(st["ms_free"], 3), "ms_partition": round(st["ms_partition"], 3), "parity": ok, "jumps": st["chain_jumps"]}))
else:
    import __graft_entry__ as g
    g.build()
    for c in sys.argv[1:] or ["65536", "32768", "16384", "8192", "4096", "2048"]:
        chunk, _, segs = c.partition(":")
        env = dict(os.environ, ISL_PIPE_CHUNK=chunk)
        if segs:
            env["ISL_PIPE_SEGMENTS"] = segs
        print(subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True).stdout.strip()[-400:])
