#!/bin/bash
# conv Cin=96: two zero-filled 64-channel chunks (128-byte rows) vs three 32-channel chunks per stage vs one
mkdir -p gpurun_out
L=gpurun_out/r2_conv_pad64_ab.log
: > $L
timeout 900 python -m pytest tests/test_gpu_vae.py tests/test_gpu_baseline_width.py -x -q -m gpu 2>&1 | tail -5 | tee -a $L
for V in "B200_CONV_PAD64=0 B200_CONV_KC=1" "B200_CONV_PAD64=0 B200_CONV_KC=3" "B200_CONV_PAD64=1" "B200_CONV_PAD64=0 B200_CONV_KC=3" "B200_CONV_PAD64=1" "B200_CONV_PAD64=1 B200_CONV_TB=4"; do
  echo "== $V :: scripts/vae_one_tile.py 5" | tee -a $L
  env $V timeout 300 python scripts/vae_one_tile.py 5 2>&1 | tail -1 | tee -a $L
done
