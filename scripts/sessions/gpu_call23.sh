#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call23.log
: > $L
timeout 600 python -m pytest tests/test_gpu_hy15_vae.py tests/test_gpu_vae.py -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
timeout 200 python scripts/row_kernels_bw.py > gpurun_out/row_kernels_bw.json 2>> $L; echo "rc=$?" >> $L
timeout 200 python scripts/bench_hy15_vae.py --tiles 16 >> $L 2>&1; echo "rc=$?" >> $L
tail -c 2500 $L
