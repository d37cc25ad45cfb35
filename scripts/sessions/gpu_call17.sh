#!/bin/bash
# N=4: dual-stream sharding check + HunyuanVideo-1.5 full-size CFG step (cfg2 x sp2)
mkdir -p gpurun_out
L=gpurun_out/call17_n4.log
: > $L
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 \
   scripts/gpu_mp_check_mmdit.py >> $L 2>&1; echo "rc=$?" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 \
   scripts/bench_hy15.py --steps 2 --warmup 1 > gpurun_out/bench_hy15_n4.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_hy15_n4.json >> $L
tail -c 3500 $L
