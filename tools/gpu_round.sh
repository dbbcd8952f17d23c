#!/bin/bash
# One GPU-box visit: tests, smoke, every bench config (own + reference arms).  Outputs under gpurun_out/<tag>_*.
tag=${1:-r2}
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt; lscpu | grep "Model name" >> gpurun_out/${tag}_host.txt; nvidia-smi -L >> gpurun_out/${tag}_host.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${tag}_smoke.log
for c in ${CONFIGS:-c4 c1 c2 c3 c3bf c5}; do
  timeout 600 python bench.py --config $c > gpurun_out/${tag}_bench_$c.json 2> gpurun_out/${tag}_bench_$c.err; echo "bench $c rc=$?"
done
for c in ${REFCONFIGS:-c4}; do
  timeout 600 python bench.py --config $c --impl reference > gpurun_out/${tag}_ref_$c.json 2> gpurun_out/${tag}_ref_$c.err; echo "ref $c rc=$?"
done
tail -3 gpurun_out/${tag}_tests.log
