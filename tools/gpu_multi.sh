#!/bin/bash
# Multi-GPU visit: bench c4 at N ranks (own arm), outputs under gpurun_out/<tag>_n<N>.*
tag=${1:-r2}; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/${tag}_n${N}.json 2> gpurun_out/${tag}_n${N}.err; echo "bench n=$N rc=$?"
grep -c "NCCL INFO" gpurun_out/${tag}_n${N}.err; grep -m3 "nranks\|NVLS\|P2P" gpurun_out/${tag}_n${N}.err | head -5
tail -c 600 gpurun_out/${tag}_n${N}.json
