#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 scripts/gpu_mp_check_mmdit.py > gpurun_out/r2_call37.log 2>&1
grep -v "^W0\|^\[W\|OMP_NUM" gpurun_out/r2_call37.log | tail -40
