#!/bin/bash
# One GPU-box visit for the profiles of a round: ncu launch list of the bench command, one --set full capture of the dominant kernel,
# per-cell chain costs.  Outputs under gpurun_out/<tag>_*.  (Under ncu kernels are serialised: the engine switches the fed / open
# streams off by itself; numbers printed by a run under ncu are never bench values.)
tag=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 > gpurun_out/${tag}_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pipeline -s 3 -c 1 -o gpurun_out/${tag}_pipeline \
    python bench.py --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "set full rc=$?"
timeout 300 python tools/chain_cost.py > gpurun_out/${tag}_chain_cost.txt 2>&1; echo "chain cost rc=$?"
ls -la gpurun_out/ | grep ${tag}
