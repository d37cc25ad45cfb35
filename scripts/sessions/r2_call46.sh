#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_vae_midattn_batched.log
: > $L
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_baseline_width.py -x -q -m gpu 2>&1 | tail -4 | tee -a $L
timeout 300 python scripts/vae_one_tile.py 5 2>&1 | tail -1 | tee -a $L
timeout 300 python scripts/vae_one_tile.py 5 2>&1 | tail -1 | tee -a $L
