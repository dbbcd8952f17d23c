#!/bin/bash
# first GPU visit of the speculative rounds: its tests, then bench c4 / c3 with the rounds forced on and off
tag=${1:-r2s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_spec.py -x -q > gpurun_out/${tag}_tests_spec.log 2>&1; echo "spec tests rc=$?"; tail -15 gpurun_out/${tag}_tests_spec.log
for sp in 1 0; do
  for c in c4 c3; do ISL_SPEC=$sp timeout 600 python bench.py --config $c --faithful-ops 500 > gpurun_out/${tag}_bench_${c}_spec$sp.json 2> gpurun_out/${tag}_bench_${c}_spec$sp.err; echo "bench $c spec=$sp rc=$?"; done
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2s_bench_*_spec*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    except Exception as e:
        print(f, 'no line', e); continue
    sc=d.get('strict_causal',{})
    print(f, 'value %.1f M/s  ms %.3f  e2e %.1f  parity %s | strict dev %.1f e2e %.1f perbatch %.1f | replay %.1f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['parity'][:9], sc.get('value',0)/1e6, sc.get('e2e_value',0)/1e6, sc.get('per_batch_calls_value',0)/1e6, d.get('replay_value',0)/1e6))
PY
