#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/call11.log
: > $L
timeout 600 python -m pytest tests/test_gpu_hy15.py -m gpu -q >> $L 2>&1; echo "rc=$?" >> $L
tail -c 3000 $L
