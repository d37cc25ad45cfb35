#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call31_n2.log
: > $L
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/gpu_mp_check.py >> $L 2>&1; echo "rc=$?" >> $L
tail -c 1800 $L
