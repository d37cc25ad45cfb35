#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_conv_out_tb.log
: > $L
for V in "B200_CONV_TB=0" "B200_CONV_TB=4" "B200_CONV_TB=2" "B200_CONV_TB=0 B200_CONV_KC=1"; do
  echo "== $V" | tee -a $L
  env $V python scripts/conv_one.py 81 256 256 96 16 2>&1 | tail -1 | tee -a $L
done
