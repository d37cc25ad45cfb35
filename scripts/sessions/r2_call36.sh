#!/bin/bash
# 2 GPUs: joint (dual-stream) fused peer-memory exchange: one-GPU addressing test, 2-GPU bit-identity, Wan multi-GPU parity
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "scatter" 2>&1 | tail -3 | tee gpurun_out/r2_call36.log
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -8 | tee -a gpurun_out/r2_call36.log
