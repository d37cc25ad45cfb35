#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_conv_one.log
: > $L
for V in "B200_CONV_TB=0 B200_CONV_KC=1" "B200_CONV_TB=0" "B200_CONV_TB=4" "B200_CONV_TB=0 B200_CONV_PAD64=1"; do
  echo "== $V" | tee -a $L
  env $V python scripts/conv_one.py 2>&1 | tail -1 | tee -a $L
  env $V python scripts/conv_one.py 81 128 128 192 192 2>&1 | tail -1 | tee -a $L
done
B200_CONV_TB=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3d -s 2 -c 1 -f -o gpurun_out/r2_conv96_plain python scripts/conv_one.py 81 256 256 96 96 1 2>&1 | tail -1
B200_CONV_TB=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3d -s 2 -c 1 -f -o gpurun_out/r2_conv96_tb python scripts/conv_one.py 81 256 256 96 96 1 2>&1 | tail -1
