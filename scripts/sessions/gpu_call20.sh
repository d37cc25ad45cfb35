#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call20.log
: > $L
timeout 300 python -m pytest tests/test_gpu_hy15_vae.py -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
timeout 300 python scripts/gemm_shapes.py > gpurun_out/gemm_shapes.json 2>> $L; echo "rc=$?" >> $L
tail -c 3500 $L
