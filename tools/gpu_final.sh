#!/bin/bash
# Final GPU visit of a round on one B200: every GPU test, every bench config (+ the reference arm), the ncu launch list of the bench
# command, one ncu --set full capture of the dominant kernel (k_pipeline resolving config 4 with one batch in flight), the round trace.
tag=${1:-r02}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${tag}_tests.log
for c in c4 c3 c2 c1 c3bf c5; do timeout 900 python bench.py --config $c > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; echo "bench $c rc=$?"; done
timeout 600 python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; echo "reference arm rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pipeline -s 2 -c 1 -o gpurun_out/${tag}_pipeline \
    python tools/profile_c4_stream.py 1 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "set full rc=$?"
timeout 300 python tools/spec_trace.py 2 > gpurun_out/${tag}_spec_trace.txt 2>&1; echo "spec trace rc=$?"
ls -la gpurun_out/ | grep ${tag}_ | head -30
