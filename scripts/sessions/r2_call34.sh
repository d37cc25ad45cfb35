#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_conv_vec32_ab.log
: > $L
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_hy15_vae.py tests/test_gpu_baseline_width.py -x -q -m gpu 2>&1 | tail -3 | tee -a $L
for V in "B200_CONV_TB=0" "B200_CONV_TB=4" "B200_CONV_TB=2"; do
  echo "== $V" | tee -a $L
  env $V python scripts/conv_one.py 2>&1 | tail -1 | tee -a $L
  env $V python scripts/conv_one.py 81 128 128 192 192 2>&1 | tail -1 | tee -a $L
  env $V python scripts/conv_one.py 41 64 64 384 384 2>&1 | tail -1 | tee -a $L
  env $V timeout 300 python scripts/vae_one_tile.py 5 2>&1 | tail -1 | tee -a $L
done
