#!/bin/bash
# GPU visit: full tests on the default build, A/B of the decision loop, ncu --set full of the stream kernel, bench c4 + c3 + c2
tag=${1:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/${tag}_tests.log
bash tools/gpu_ab.sh ${tag} v0 vu vud
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pipeline -s 2 -c 1 -o gpurun_out/r02_pipeline \
    python tools/profile_c4_stream.py > gpurun_out/${tag}_ncu_full.log 2>&1; echo "set full rc=$?"
for c in c4 c3 c2; do timeout 600 python bench.py --config $c > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; echo "bench $c rc=$?"; done
