#!/bin/bash
# conv + RMS-norm + SiLU fused epilogue: tests, then A/B on one tile and the full decode
mkdir -p gpurun_out
L=gpurun_out/r2_conv_norm_fuse_ab.log
: > $L
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_baseline_width.py -x -q -m gpu 2>&1 | tail -5 | tee -a $L
for V in "B200_VAE_FUSE_NORM=0" "B200_VAE_FUSE_NORM=1" "B200_VAE_FUSE_NORM=0" "B200_VAE_FUSE_NORM=1"; do
  echo "== $V :: scripts/vae_one_tile.py 5" | tee -a $L
  env $V timeout 300 python scripts/vae_one_tile.py 5 2>&1 | tail -1 | tee -a $L
done
