#!/bin/bash
# HunyuanVideo-1.5 720p x 129f CFG step on 8 GPUs (cfg2 x sp4): fused peer-memory exchange vs NCCL all-to-all, same box
mkdir -p gpurun_out
for V in "--p2p" ""; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 scripts/bench_hy15.py --steps 2 --warmup 1 $V > gpurun_out/r2_bench_hy15_n8${V/--/_}.json 2> gpurun_out/r2_bench_hy15_n8${V/--/_}.err
  tail -c 1200 gpurun_out/r2_bench_hy15_n8${V/--/_}.json
done
