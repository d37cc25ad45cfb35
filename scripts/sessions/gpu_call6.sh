#!/bin/bash
# N=8 bench (fused exchange), bounded
mkdir -p gpurun_out
L=gpurun_out/call6_n8.log
: > $L
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 \
   bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_n8.json 2>> $L
cat gpurun_out/bench_n8.json >> $L
tail -c 2500 $L
