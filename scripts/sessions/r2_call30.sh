#!/bin/bash
# conv: all 96 channels of a tap per stage (KC=3) + descriptor increments; parity tests, then A/B on one full-size VAE tile
mkdir -p gpurun_out
L=gpurun_out/r2_conv_kc_ab.log
: > $L
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_hy15_vae.py tests/test_gpu_baseline_width.py -x -q -m gpu 2>&1 | tail -5 | tee -a $L
for KC in 1 3 1 3; do
  echo "== B200_CONV_KC=$KC :: scripts/vae_one_tile.py 5" | tee -a $L
  B200_CONV_KC=$KC timeout 300 python scripts/vae_one_tile.py 5 2>&1 | tail -1 | tee -a $L
done
echo "== B200_CONV_TB=4 :: scripts/vae_one_tile.py 5" | tee -a $L
B200_CONV_TB=4 timeout 300 python scripts/vae_one_tile.py 5 2>&1 | tail -1 | tee -a $L
