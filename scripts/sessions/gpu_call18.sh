#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call18_n8.log
: > $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 \
   scripts/bench_hy15.py --steps 2 --warmup 1 > gpurun_out/bench_hy15_n8.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_hy15_n8.json >> $L
tail -c 2500 $L
