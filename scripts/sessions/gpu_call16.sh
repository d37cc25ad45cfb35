#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call16_n${1:-2}.log
: > $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${1:-2} --master-addr 127.0.0.1 --master-port 29512 \
   scripts/gpu_mp_check_mmdit.py >> $L 2>&1; echo "rc=$?" >> $L
tail -c 3500 $L
