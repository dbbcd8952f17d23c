#!/bin/bash
# GPU visit: every GPU test, bench c4 (new headline: strict causal) and c3
tag=${1:-r2u}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/${tag}_tests.log
for c in c4 c3; do timeout 900 python bench.py --config $c > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; echo "bench $c rc=$?"; tail -2 gpurun_out/${tag}_bench_$c.err; done
python - <<PY
import json
for c in ("c4","c3"):
    try: d=json.loads([l for l in open("gpurun_out/${tag}_bench_%s.json"%c) if l.startswith("{")][-1])
    except Exception as e: print(c,"no line",e); continue
    print(c,"value %.1f M/s ms %.3f e2e %.1f M/s parity %s"%(d["value"]/1e6,d["ms_per_step"],d["e2e"]["value"]/1e6,d["parity"][:9]))
    for k in ("strict_causal","causal_feed"):
        if k in d: print("  ",k,{x:(round(v/1e6,1) if isinstance(v,float) and v>1e5 else v) for x,v in d[k].items() if x!="workload"})
    if "replay_value" in d: print("   replay %.1f e2e %.1f"%(d["replay_value"]/1e6,d["replay_e2e_value"]/1e6))
PY
