#!/bin/bash
# temporally blocked conv (conv3d_tb_kernel): parity tests, then A/B of one full-size VAE tile: TB off / 2 / 4
mkdir -p gpurun_out
L=gpurun_out/r2_conv_tb_ab.log
: > $L
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_hy15_vae.py tests/test_gpu_baseline_width.py -x -q -m gpu 2>&1 | tail -5 | tee -a $L
for TB in 0 4 2 0 4; do
  echo "== B200_CONV_TB=$TB :: scripts/vae_one_tile.py 5" | tee -a $L
  B200_CONV_TB=$TB timeout 300 python scripts/vae_one_tile.py 5 2>&1 | tail -1 | tee -a $L
done
