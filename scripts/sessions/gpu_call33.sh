#!/bin/bash
mkdir -p gpurun_out
timeout 75 python -m pytest tests/test_gpu_variants.py -m gpu -q -x > gpurun_out/call33.log 2>&1; echo "rc=$?" >> gpurun_out/call33.log
tail -c 1500 gpurun_out/call33.log
