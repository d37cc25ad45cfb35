#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call14.log
: > $L
timeout 600 python -m pytest tests/test_gpu_hy15_vae.py tests/test_gpu_vae.py -m gpu -q --durations=5 >> $L 2>&1; echo "rc=$?" >> $L
echo "== bench hy15 vae (8 tiles)" >> $L
timeout 300 python scripts/bench_hy15_vae.py --tiles 8 > gpurun_out/bench_hy15_vae.json 2>> $L; echo "rc=$?" >> $L
cat gpurun_out/bench_hy15_vae.json >> $L
tail -c 6000 $L
