#!/bin/bash
# temporally blocked conv v2 (KC-merged stages, weight buffer per (kh,kw), per-accumulator release, hoisted epilogue loads)
mkdir -p gpurun_out
L=gpurun_out/r2_conv_tb2_ab.log
: > $L
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_hy15_vae.py tests/test_gpu_baseline_width.py -x -q -m gpu 2>&1 | tail -5 | tee -a $L
B200_CONV_TB=4 timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_hy15_vae.py -x -q -m gpu 2>&1 | tail -3 | tee -a $L
for V in "B200_CONV_TB=0" "B200_CONV_TB=-1" "B200_CONV_TB=2" "B200_CONV_TB=4" "B200_CONV_TB=0" "B200_CONV_TB=-1"; do
  echo "== $V :: scripts/vae_one_tile.py 5" | tee -a $L
  env $V timeout 300 python scripts/vae_one_tile.py 5 2>&1 | tail -1 | tee -a $L
done
